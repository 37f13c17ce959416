"""The engine's host thread pool (dabstar_b200/csrc/hostpool.h): compiled into a small test program with g++ (no CUDA needed)
and hammered with parallel_for calls of every size, nested data dependencies between regions, pools of 0..7 workers, polling and
sleeping workers and wake-ups ahead of a region (prewake)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include "hostpool.h"
#include <cstdio>
#include <numeric>
int main()
{
  for (int workers = 0; workers <= 7; workers++)
  {
    dab::HostPool pool(workers, workers % 2 ? 1000 : 0); // workers that poll for 1 ms, and workers that sleep at once
    std::vector<long long> v(5000);
    long long want = 0;
    for (int round = 0; round < 400; round++)
    {
      const int n = round % 97 == 0 ? 5000 : (round * 37) % 131;   // incl. 0 and 1
      if (round % 50 == 49) std::this_thread::sleep_for(std::chrono::milliseconds(3)); // let the workers fall asleep
      if (round % 3 == 0) pool.prewake();                                              // with and without a region behind it
      if (round % 100 == 99) { pool.prewake(); std::this_thread::sleep_for(std::chrono::milliseconds(3)); } // a wake-up nothing follows
      // region A writes, region B reads what A wrote (a region is complete when parallel_for returns)
      pool.parallel_for(n, [&](int i) { v[(size_t)i] = (long long)i * round; });
      std::atomic<long long> sum{ 0 };
      pool.parallel_for(n, [&](int i) { sum.fetch_add(v[(size_t)i], std::memory_order_relaxed); });
      want = (long long)round * ((long long)n * (n - 1) / 2);
      if (sum.load() != want) { printf("mismatch workers=%d round=%d n=%d: %lld != %lld\n", workers, round, n, sum.load(), want); return 1; }
    }
  }
  printf("ok %d\n", dab::HostPool::default_workers());
  return 0;
}
"""


def test_hostpool_parallel_for(tmp_path):
    src = tmp_path / "pool_test.cpp"
    src.write_text(SRC)
    exe = tmp_path / "pool_test"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "dabstar_b200", "csrc"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
