// tii_kernels.cu — the reference's TII detector (ofdm/tii_detector.cpp) for a batch of recordings: one CTA per detector.
// Null-symbol spectra are accumulated (add_to_tii_buffer, :154-160), neighbouring carriers are multiplied into the 768
// pair products and low-pass filtered (_decode_and_accumulate_carrier_pairs, :247-267), the four frequency blocks are
// collapsed with and without the "non-ETSI" phase correction (:317-343), and every comb (sub id) is searched for four or
// more groups above the threshold (:162-241, :345-472). Products and sums use unfused multiplies and adds in the
// reference's order, so the identifications equal the CPU's; the result list is sorted by strength on the host.
#include "kernels.h"

namespace dab
{
namespace
{
constexpr int TII_THREADS = 192; // cBlockSize192: one thread per collapsed carrier pair
constexpr int TII_BLOCKS = 4, TII_GROUPS = 8, TII_GROUP = 24, TII_PAIRS = 768, TII_MAIN = 70;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float cabs_ref(float2 v) { return sqrtf(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))); }
__device__ __forceinline__ float2 turn_phase(float2 v, int phase) // tii_detector.cpp:295-309
{
  switch (phase)
  {
  case 3: return make_float2(-v.y, v.x);
  case 2: return make_float2(-v.x, -v.y);
  case 1: return make_float2(v.y, -v.x);
  default: return v;
  }
}

// null_sum[d][bin] += fft[d][0][bin] + fft[d][1][bin] + ... in that order
__global__ void __launch_bounds__(256) k_tii_add(const float2 * __restrict__ fft, int n_symbols, float2 * __restrict__ null_sum)
{
  const int d = blockIdx.y, bin = blockIdx.x * 256 + threadIdx.x;
  float2 acc = null_sum[(size_t)d * T_U + bin];
  const float2 * src = fft + (size_t)d * n_symbols * T_U + bin;
  for (int s = 0; s < n_symbols; s++) acc = cadd(acc, src[(size_t)s * T_U]);
  null_sum[(size_t)d * T_U + bin] = acc;
}

__global__ void __launch_bounds__(TII_THREADS) k_tii_process(float2 * __restrict__ null_sum, float2 * __restrict__ decoded, const uint8_t * __restrict__ pattern,
                                                             const uint8_t * __restrict__ phase_corr, float threshold_factor, int collisions, int sub_id_coll,
                                                             TiiResultDev * __restrict__ out, int cap, int * __restrict__ counts)
{
  __shared__ float2 etsi_c[TII_THREADS], non_c[TII_THREADS];
  __shared__ float etsi_f[TII_THREADS], non_f[TII_THREADS];
  __shared__ float red[TII_THREADS / 32];
  __shared__ float s_avg[TII_GROUP];
  __shared__ int n_out;
  const int d = blockIdx.x, i = threadIdx.x;
  float2 * ns = null_sum + (size_t)d * T_U;
  float2 * dec = decoded + (size_t)d * TII_PAIRS;
  if (i == 0) n_out = 0;

  // pair products, IIR (alpha 0.01), single-carrier suppression, collapse: everything a thread needs is its own four pairs
  float2 buf[TII_BLOCKS];
  float mag[TII_BLOCKS];
#pragma unroll
  for (int b = 0; b < TII_BLOCKS; b++)
  {
    const int p = i + b * TII_THREADS, k = -K_CARR / 2 + 2 * p, f = k < 0 ? k + T_U : k + 1;
    const float2 x = ns[f], y = ns[f + 1];
    const float2 prod = make_float2(__fadd_rn(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y)), __fsub_rn(__fmul_rn(x.y, y.x), __fmul_rn(x.x, y.y)));
    float2 v = dec[p];
    v.x = __fadd_rn(v.x, __fmul_rn(0.01f, __fsub_rn(prod.x, v.x)));
    v.y = __fadd_rn(v.y, __fmul_rn(0.01f, __fsub_rn(prod.y, v.y)));
    dec[p] = v;
    buf[b] = v;
    mag[b] = cabs_ref(v);
  }
  {
    // _remove_single_carrier_values (:269-293), mCarrierDelete = true
    float mx = 0.0f, sum = 0.0f;
    int index = 0;
#pragma unroll
    for (int b = 0; b < TII_BLOCKS; b++)
    {
      sum = __fadd_rn(sum, mag[b]);
      if (mag[b] > mx) { mx = mag[b]; index = b; }
    }
    const float mn = __fdiv_rn(__fsub_rn(sum, mx), (float)(TII_BLOCKS - 1));
    if (sum < __fmul_rn(mx, 1.5f) && mx > 0.0f)
    {
      const float sc = __fdiv_rn(mn, mx);
#pragma unroll
      for (int b = 0; b < TII_BLOCKS; b++)
        if (b == index) buf[b] = make_float2(__fmul_rn(buf[b].x, sc), __fmul_rn(buf[b].y, sc));
    }
  }
  float2 e = make_float2(0.f, 0.f), n = make_float2(0.f, 0.f);
#pragma unroll
  for (int b = 0; b < TII_BLOCKS; b++)
  {
    e = cadd(e, buf[b]);
    n = cadd(n, turn_phase(buf[b], phase_corr[i + b * TII_THREADS]));
  }
  etsi_c[i] = e;
  non_c[i] = n;
  const float ef = cabs_ref(e), nf = cabs_ref(n);
  etsi_f[i] = ef;
  non_f[i] = nf;
  // strongest collapsed carrier (a maximum does not depend on the order)
  float m = fmaxf(ef, nf);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((i & 31) == 0) red[i >> 5] = m;
  // the accumulated spectra are used up (_reset_null_symbol_buffer)
  for (int f = i; f < T_U; f += TII_THREADS) ns[f] = make_float2(0.f, 0.f);
  __syncthreads();
  float max_abs = 0.0f;
#pragma unroll
  for (int w = 0; w < TII_THREADS / 32; w++) max_abs = fmaxf(max_abs, red[w]);
  // _calculate_average_noise (:485-500): the weakest comb
  if (i < TII_GROUP)
  {
    float avg = 0.0f;
    for (int g = 0; g < TII_GROUPS; g++) avg = __fadd_rn(avg, etsi_f[i + g * TII_GROUP]);
    s_avg[i] = __fdiv_rn(avg, (float)TII_GROUPS);
  }
  __syncthreads();
  if (i < TII_GROUP)
  {
    float noise = 1e9f;
    for (int s = 0; s < TII_GROUP; s++) noise = fminf(noise, s_avg[s]);
    const float thr = __fmul_rn(noise, threshold_factor);
    const int sub = i;

    // _compare_etsi_and_non_etsi (:385-437)
    float2 es = make_float2(0.f, 0.f), nsum = make_float2(0.f, 0.f);
    int ec = 0, nc = 0;
    unsigned ep = 0, np = 0;
    for (int g = 0; g < TII_GROUPS; g++)
    {
      const int idx = sub + g * TII_GROUP;
      if (etsi_f[idx] > thr) { ec++; ep |= 0x80u >> g; es = cadd(es, etsi_c[idx]); }
      if (non_f[idx] > thr) { nc++; np |= 0x80u >> g; nsum = cadd(nsum, non_c[idx]); }
    }
    bool non_etsi = false;
    if (ec >= 4 || nc >= 4) non_etsi = cabs_ref(nsum) > cabs_ref(es);
    float2 sum = non_etsi ? nsum : es;
    const int count = non_etsi ? nc : ec;
    const unsigned pat = non_etsi ? np : ep;
    const float2 * ct = non_etsi ? non_c : etsi_c;
    const float * ft = non_etsi ? non_f : etsi_f;
    if (count >= 4)
    {
      int main_id = 0;
      if (count == 4)
      {
        for (int mm = 0; mm < TII_MAIN; mm++) if (pattern[mm] == pat) { main_id = mm; break; } // _find_exact_main_id_match (:345-357)
      }
      else
      {
        // _find_best_main_id_match (:359-383)
        float max_level = 0.0f;
        main_id = -1;
        sum = make_float2(0.f, 0.f);
        for (int mm = 0; mm < TII_MAIN; mm++)
        {
          float2 val = make_float2(0.f, 0.f);
          for (int g = 0; g < TII_GROUPS; g++)
            if (pattern[mm] & (0x80u >> g)) val = cadd(val, ct[sub + TII_GROUP * g]);
          const float a = cabs_ref(val);
          if (a > max_level) { max_level = a; sum = val; main_id = mm; }
        }
      }
      auto emit = [&](int mid, float2 s, float div) {
        const int at = atomicAdd(&n_out, 1);
        if (at < cap)
          out[(size_t)d * cap + at] = TiiResultDev{ mid & 0xff, sub, __fdiv_rn(__fdiv_rn(cabs_ref(s), max_abs), div), atan2f(s.y, s.x) * 57.29577951308232f, non_etsi ? 1 : 0 };
      };
      emit(main_id, sum, 4.0f);
      if (count > 4 && collisions)
      {
        // _find_collisions (:439-472)
        float2 s2 = make_float2(0.f, 0.f);
        for (int g = 0; g < TII_GROUPS; g++)
          if (!(pattern[main_id & 0xff] & (0x80u >> g)))
          {
            const int idx = sub + TII_GROUP * g;
            if (ft[idx] > thr) s2 = cadd(s2, ct[idx]);
          }
        if (sub == sub_id_coll)
        {
          for (int mm = 0; mm < TII_MAIN; mm++)
            if (__popc(pattern[mm] & pat) == 4 && mm != main_id) emit(mm, s2, (float)(count - 4));
        }
        else emit(99, s2, (float)(count - 4));
      }
    }
  }
  __syncthreads();
  if (i == 0) counts[d] = n_out; // may exceed cap: the caller then knows the list was cut
}

} // namespace

cudaError_t launch_tii_add(cudaStream_t s, const float2 * fft, int n_detectors, int n_symbols, float2 * null_sum, unsigned long long * lc)
{
  if (n_detectors <= 0 || n_symbols <= 0) return cudaSuccess;
  k_tii_add<<<dim3(T_U / 256, (unsigned)n_detectors), 256, 0, s>>>(fft, n_symbols, null_sum);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_tii_process(cudaStream_t s, int n_detectors, float2 * null_sum, float2 * decoded, const uint8_t * pattern, const uint8_t * phase_corr,
                               float threshold_factor, int collisions, int sub_id_coll, TiiResultDev * out, int cap, int * counts, unsigned long long * lc)
{
  if (n_detectors <= 0) return cudaSuccess;
  k_tii_process<<<(unsigned)n_detectors, TII_THREADS, 0, s>>>(null_sum, decoded, pattern, phase_corr, threshold_factor, collisions, sub_id_coll, out, cap, counts);
  if (lc) (*lc)++;
  return cudaGetLastError();
}
} // namespace dab
