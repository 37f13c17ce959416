// dciq_kernels.cu — SampleReader's optional DC removal and IQ-imbalance correction (ofdm/sample_reader.cpp:216-243,
// default off, switched by set_dc_and_iq_correction) as a pre-pass over a whole recording.
//
// The reference runs five first-order IIRs sample by sample, all with alpha = 1 / 2 048 000:
//   meanI, meanQ  <- v_i, v_q                       x_i = v_i - meanI, x_q = v_q - meanQ
//   meanII, meanIQ <- x_i^2, x_i x_q                phi = meanIQ / meanII, x_qc = x_q - phi x_i
//   meanQQ <- x_qc^2                                gainQ = sqrt(meanII / meanQQ), out = (x_i, x_qc gainQ)
// Each filter y <- y + alpha (x - y) is the affine map y -> w y + alpha x (w = 1 - alpha), so a level is a scan of affine
// maps: chunks of 1024 samples are scanned inside a CTA, the chunk totals are chained by one thread, and the next level
// starts from the values the previous one produced (three levels, each needs the one before per sample). The scan runs in
// double precision: 1 - alpha is not a float, and a float w would change the time constant by 2 %. The result therefore
// equals the recurrences in exact arithmetic; the reference's own float rounding (a random walk of half-ulp errors over
// the 2 M-sample memory of the filters) is the remaining difference, about 1e-6 of the sample amplitude.
#include "kernels.h"

namespace dab
{
namespace
{
constexpr int DCIQ_THREADS = 256;
constexpr int DCIQ_PER_THREAD = 4;
constexpr int DCIQ_CHUNK = DCIQ_THREADS * DCIQ_PER_THREAD;

struct Aff2 // y -> a y + (b0, b1) for two filters with the same coefficient
{
  double a, b0, b1;
};
__device__ __forceinline__ Aff2 then(const Aff2 & first, const Aff2 & second) // second o first
{
  return Aff2{ second.a * first.a, second.a * first.b0 + second.b0, second.a * first.b1 + second.b1 };
}
__device__ __forceinline__ double shfl_up_d(double v, int d)
{
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, d);
  hi = __shfl_up_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}

// Exclusive scan of one affine map per thread over the CTA; returns the map of everything before this thread and, in
// total, the map of the whole CTA. smem: DCIQ_THREADS / 32 entries.
__device__ Aff2 block_scan(Aff2 mine, Aff2 * smem, Aff2 & total)
{
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  Aff2 inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    Aff2 o{ shfl_up_d(inc.a, d), shfl_up_d(inc.b0, d), shfl_up_d(inc.b1, d) };
    if (lane >= d) inc = then(o, inc);
  }
  if (lane == 31) smem[wp] = inc;
  __syncthreads();
  Aff2 before{ 1.0, 0.0, 0.0 };
  for (int i = 0; i < wp; i++) before = then(before, smem[i]);
  total = before;
  for (int i = wp; i < DCIQ_THREADS / 32; i++) total = then(total, smem[i]);
  __syncthreads();
  // exclusive prefix inside the warp: the inclusive one of the lane below
  Aff2 excl{ shfl_up_d(inc.a, 1), shfl_up_d(inc.b0, 1), shfl_up_d(inc.b1, 1) };
  if (lane == 0) excl = Aff2{ 1.0, 0.0, 0.0 };
  return then(before, excl);
}

// One level of the cascade for one chunk per CTA.
//   LEVEL 1: totals of (meanI, meanQ)                              -> agg[0]
//   LEVEL 2: apply level 1; DO_IQ ? totals of (meanII, meanIQ) -> agg[1] : write v - mean
//   LEVEL 3: apply levels 1, 2; totals of (meanQQ, -)              -> agg[2]
//   LEVEL 4: apply levels 1..3; write (x_i, x_qc gainQ)
// start[l][c]: filter values of level l + 1 at the start of chunk c.
template <int LEVEL, bool DO_IQ>
__global__ void __launch_bounds__(DCIQ_THREADS) k_dciq(const float2 * in, long long n, double alpha, const double2 * __restrict__ start1,
                                                       const double2 * __restrict__ start2, const double2 * __restrict__ start3, Aff2 * __restrict__ agg,
                                                       float2 * out) // (in == out is allowed: a CTA reads its chunk before it writes it)
{
  __shared__ Aff2 smem[DCIQ_THREADS / 32];
  const long long c = blockIdx.x;
  const long long base = c * DCIQ_CHUNK + (long long)threadIdx.x * DCIQ_PER_THREAD;
  const double w = 1.0 - alpha;
  double vi[DCIQ_PER_THREAD], vq[DCIQ_PER_THREAD];
  bool ok[DCIQ_PER_THREAD];
#pragma unroll
  for (int j = 0; j < DCIQ_PER_THREAD; j++)
  {
    ok[j] = base + j < n;
    const float2 v = ok[j] ? in[base + j] : make_float2(0.f, 0.f);
    vi[j] = (double)v.x;
    vq[j] = (double)v.y;
  }
  // ---- level 1
  Aff2 m{ 1.0, 0.0, 0.0 };
#pragma unroll
  for (int j = 0; j < DCIQ_PER_THREAD; j++)
    if (ok[j]) m = then(m, Aff2{ w, alpha * vi[j], alpha * vq[j] });
  Aff2 tot;
  Aff2 pre = block_scan(m, smem, tot);
  if (LEVEL == 1) { if (threadIdx.x == 0) agg[c] = tot; return; }
  double xi[DCIQ_PER_THREAD], xq[DCIQ_PER_THREAD];
  {
    const double2 s = start1[c];
    double mi = pre.a * s.x + pre.b0, mq = pre.a * s.y + pre.b1;
#pragma unroll
    for (int j = 0; j < DCIQ_PER_THREAD; j++)
    {
      if (ok[j]) { mi = w * mi + alpha * vi[j]; mq = w * mq + alpha * vq[j]; }
      xi[j] = vi[j] - mi;
      xq[j] = vq[j] - mq;
    }
  }
  if (!DO_IQ)
  {
#pragma unroll
    for (int j = 0; j < DCIQ_PER_THREAD; j++)
      if (ok[j]) out[base + j] = make_float2((float)xi[j], (float)xq[j]);
    return;
  }
  // ---- level 2
  m = Aff2{ 1.0, 0.0, 0.0 };
#pragma unroll
  for (int j = 0; j < DCIQ_PER_THREAD; j++)
    if (ok[j]) m = then(m, Aff2{ w, alpha * xi[j] * xi[j], alpha * xi[j] * xq[j] });
  pre = block_scan(m, smem, tot);
  if (LEVEL == 2) { if (threadIdx.x == 0) agg[c] = tot; return; }
  double mii[DCIQ_PER_THREAD], xqc[DCIQ_PER_THREAD];
  {
    const double2 s = start2[c];
    double ii = pre.a * s.x + pre.b0, iq = pre.a * s.y + pre.b1;
#pragma unroll
    for (int j = 0; j < DCIQ_PER_THREAD; j++)
    {
      if (ok[j]) { ii = w * ii + alpha * xi[j] * xi[j]; iq = w * iq + alpha * xi[j] * xq[j]; }
      mii[j] = ii;
      xqc[j] = xq[j] - iq / ii * xi[j];
    }
  }
  // ---- level 3
  m = Aff2{ 1.0, 0.0, 0.0 };
#pragma unroll
  for (int j = 0; j < DCIQ_PER_THREAD; j++)
    if (ok[j]) m = then(m, Aff2{ w, alpha * xqc[j] * xqc[j], 0.0 });
  pre = block_scan(m, smem, tot);
  if (LEVEL == 3) { if (threadIdx.x == 0) agg[c] = tot; return; }
  {
    const double2 s = start3[c];
    double qq = pre.a * s.x + pre.b0;
#pragma unroll
    for (int j = 0; j < DCIQ_PER_THREAD; j++)
    {
      if (ok[j])
      {
        qq = w * qq + alpha * xqc[j] * xqc[j];
        out[base + j] = make_float2((float)xi[j], (float)(xqc[j] * sqrt(mii[j] / qq)));
      }
    }
  }
}

// Chains the chunk totals: start[c + 1] = agg[c](start[c]); start[0] = the filter values before the first sample.
__global__ void k_dciq_chain(const Aff2 * __restrict__ agg, long long n_chunks, double2 s0, double2 * __restrict__ start)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double2 s = s0;
  for (long long c = 0; c < n_chunks; c++)
  {
    start[c] = s;
    const Aff2 a = agg[c];
    s = make_double2(a.a * s.x + a.b0, a.a * s.y + a.b1);
  }
  start[n_chunks] = s;
}
} // namespace

size_t dciq_workspace_bytes(long long n)
{
  const long long chunks = (n + DCIQ_CHUNK - 1) / DCIQ_CHUNK;
  return (size_t)chunks * sizeof(Aff2) + 3 * (size_t)(chunks + 1) * sizeof(double2) + 256;
}

// state: { meanI, meanQ, meanII, meanIQ, meanQQ } before the first sample; updated to the values after the last one.
cudaError_t launch_dc_iq_correct(cudaStream_t s, const float2 * in, long long n, bool do_iq, double alpha, double state[5], void * workspace, float2 * out,
                                 unsigned long long * lc)
{
  if (n <= 0) return cudaSuccess;
  const long long chunks = (n + DCIQ_CHUNK - 1) / DCIQ_CHUNK;
  Aff2 * agg = static_cast<Aff2 *>(workspace);
  double2 * st1 = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(agg + chunks) + 15) & ~(uintptr_t)15);
  double2 * st2 = st1 + (chunks + 1);
  double2 * st3 = st2 + (chunks + 1);
  const unsigned grid = (unsigned)chunks;
  k_dciq<1, true><<<grid, DCIQ_THREADS, 0, s>>>(in, n, alpha, nullptr, nullptr, nullptr, agg, nullptr);
  k_dciq_chain<<<1, 1, 0, s>>>(agg, chunks, make_double2(state[0], state[1]), st1);
  if (lc) *lc += 2;
  if (!do_iq)
  {
    k_dciq<2, false><<<grid, DCIQ_THREADS, 0, s>>>(in, n, alpha, st1, nullptr, nullptr, nullptr, out);
    if (lc) *lc += 1;
  }
  else
  {
    k_dciq<2, true><<<grid, DCIQ_THREADS, 0, s>>>(in, n, alpha, st1, nullptr, nullptr, agg, nullptr);
    k_dciq_chain<<<1, 1, 0, s>>>(agg, chunks, make_double2(state[2], state[3]), st2);
    k_dciq<3, true><<<grid, DCIQ_THREADS, 0, s>>>(in, n, alpha, st1, st2, nullptr, agg, nullptr);
    k_dciq_chain<<<1, 1, 0, s>>>(agg, chunks, make_double2(state[4], 0.0), st3);
    k_dciq<4, true><<<grid, DCIQ_THREADS, 0, s>>>(in, n, alpha, st1, st2, st3, nullptr, out);
    if (lc) *lc += 5;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // final filter values
  double2 f1, f2 = make_double2(state[2], state[3]), f3 = make_double2(state[4], 0.0);
  e = cudaMemcpyAsync(&f1, st1 + chunks, sizeof(double2), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && do_iq) e = cudaMemcpyAsync(&f2, st2 + chunks, sizeof(double2), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && do_iq) e = cudaMemcpyAsync(&f3, st3 + chunks, sizeof(double2), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  state[0] = f1.x; state[1] = f1.y; state[2] = f2.x; state[3] = f2.y; state[4] = f3.x;
  return cudaSuccess;
}
} // namespace dab
