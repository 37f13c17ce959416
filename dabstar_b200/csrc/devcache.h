// devcache.h — per-device cache of launch properties.
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the occupancy of a kernel are properties of (device, function):
// a process may hold contexts on several GPUs (dabstar_create takes any device index) and drive them from different
// threads, so nothing here is a plain function-local static.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

namespace dab
{
struct LaunchProps
{
  cudaError_t err = cudaSuccess;
  int ctas_per_sm = 0; // resident CTAs per SM at the given block size / dynamic shared memory
  int n_sm = 0;        // multiprocessors of the device
};

// Opts `fn` in for `optin_smem` bytes of dynamic shared memory (0: no opt-in needed) on the CURRENT device, once, and
// returns its occupancy at (threads, smem) and the device's SM count.
inline LaunchProps launch_props(const void * fn, int threads, size_t smem, size_t optin_smem)
{
  static std::mutex mu;
  static std::map<std::tuple<int, const void *, int, size_t>, LaunchProps> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(dev, fn, threads, smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  LaunchProps p;
  if (optin_smem > 0) p.err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_smem);
  if (p.err == cudaSuccess) p.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p.ctas_per_sm, fn, threads, smem);
  if (p.err == cudaSuccess && cudaDeviceGetAttribute(&p.n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) p.n_sm = 0;
  if (p.n_sm <= 0) p.n_sm = 148;
  if (p.err == cudaSuccess) cache[key] = p;
  return p;
}
} // namespace dab
