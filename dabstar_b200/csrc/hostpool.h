// hostpool.h — a few persistent host threads for the engine's per-recording control loops.
//
// DabProcessor's scalar bookkeeping (frame layout, the AFC / clock recurrences of dab_processor.cpp:205-251, PRS-peak and FIC
// verification, result filing) is independent per recording; with ~100 recordings x ~100 frames per window it is the part of a
// step the GPU waits for. parallel_for hands the recordings to the pool; the caller takes part, so a pool of 0 workers is the
// plain loop.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

namespace dab
{
class HostPool
{
public:
  // spin_us: how long an idle worker polls the generation counter before it goes to sleep. The regions of a decode step are
  // less than that apart, so the workers stay awake through a step's control phases and sleep through its long kernels;
  // prewake() gets them polling again shortly before the next region is due (waking a sleeping worker costs ~50 us, which a
  // step paid several times over: 0.4 of 12.9 ms).
  explicit HostPool(int workers, int spin_us = 1000) : spin_us_(spin_us)
  {
    for (int i = 0; i < workers; i++) threads_.emplace_back([this] { loop(); });
  }
  ~HostPool()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto & t : threads_) t.join();
  }
  int workers() const { return (int)threads_.size(); }

  // sleeping workers start polling again (for spin_us): call when a parallel_for is coming up
  void prewake()
  {
    if (threads_.empty()) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      wake_++;
    }
    cv_.notify_all();
  }

  // CPUs this process may run on, minus the caller's, capped
  static int default_workers(int cap = 7)
  {
    cpu_set_t set;
    int n = 1;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    else n = (int)std::thread::hardware_concurrency();
    return std::max(0, std::min(cap, n - 1));
  }

  // fn(i) for every i in [0, n); returns when all calls have returned. fn must not throw.
  template <class F> void parallel_for(int n, F && fn)
  {
    if (n <= 0) return;
    if (threads_.empty() || n == 1)
    {
      for (int i = 0; i < n; i++) fn(i);
      return;
    }
    std::function<void(int)> f = std::ref(fn);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f;
      n_ = n;
      next_.store(0, std::memory_order_relaxed);
      done_.store(0, std::memory_order_relaxed);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    work(f, n);
    while (done_.load(std::memory_order_acquire) < n) std::this_thread::yield();
    {
      // no worker may still be looking at `f` when it goes out of scope
      std::lock_guard<std::mutex> lk(m_);
      job_ = nullptr;
    }
    while (busy_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }

private:
  void work(std::function<void(int)> & f, int n)
  {
    int did = 0;
    for (int i = next_.fetch_add(1, std::memory_order_relaxed); i < n; i = next_.fetch_add(1, std::memory_order_relaxed)) { f(i); did++; }
    if (did) done_.fetch_add(did, std::memory_order_release);
  }
  void loop()
  {
    unsigned long long seen = 0, seen_wake = 0;
    while (true)
    {
      std::function<void(int)> * f = nullptr;
      int n = 0;
      // poll the generation counter (no lock) before sleeping
      {
        const auto t0 = std::chrono::steady_clock::now();
        for (int spin = 0; gen_.load(std::memory_order_acquire) == seen; spin++)
        {
          if ((spin & 63) == 63)
          {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
            std::this_thread::yield();
          }
          else __builtin_ia32_pause();
        }
      }
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen || stop_ || wake_ != seen_wake; });
        if (stop_) return;
        seen_wake = wake_;
        if (gen_.load(std::memory_order_acquire) == seen) continue; // woken ahead of a region: poll for it
        seen = gen_.load(std::memory_order_acquire);
        f = job_;
        n = n_;
        if (f) busy_.fetch_add(1, std::memory_order_acq_rel);
      }
      if (f)
      {
        work(*f, n);
        busy_.fetch_sub(1, std::memory_order_acq_rel);
      }
    }
  }

  const int spin_us_;
  unsigned long long wake_ = 0; // under m_
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_;
  std::function<void(int)> * job_ = nullptr;
  int n_ = 0;
  std::atomic<unsigned long long> gen_{ 0 };
  bool stop_ = false;
  std::atomic<int> next_{ 0 }, done_{ 0 }, busy_{ 0 };
};
} // namespace dab
