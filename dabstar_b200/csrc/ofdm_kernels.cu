// ofdm_kernels.cu — ingest, synchronisation, FFT and DQPSK demapping kernels for sm_100a.
//
// Reference rows: I1/I2 sample conversion and integer-Hz derotation
// (devices/filereaders/raw_files/raw_reader.cpp:66-70, ofdm/sample_reader.cpp:212-282), S1 TimeSyncer
// (ofdm/timesyncer.cpp:40-90), S2/S3 PhaseReference (ofdm/phasereference.cpp:87-280), F1 the FFTW plan
// (main/dab_processor.cpp:63), F2 cyclic-prefix correlation (main/dab_processor.cpp:317-333), D0-D3
// FreqInterleaver + OfdmDecoder (ofdm/ofdm_decoder.cpp:114-355).
#include "fft2048.cuh"
#include "kernels.h"

#include <type_traits>

namespace dab
{
namespace
{
// ------------------------------------------------------------------------------------------------ ingest
// One IQ pair -> complex float. u8: (v - 127.38)/128 (raw_reader.cpp:66-70); i16: v/32768 (xml_reader.cpp:254-372).
template <int FMT> __device__ __forceinline__ float2 load_sample(const void * __restrict__ iq, long long i)
{
  if (FMT == FMT_U8)
  {
    const uchar2 v = reinterpret_cast<const uchar2 *>(iq)[i];
    return make_float2(((float)v.x - 127.38f) * (1.0f / 128.0f), ((float)v.y - 127.38f) * (1.0f / 128.0f));
  }
  else if (FMT == FMT_I16)
  {
    const short2 v = reinterpret_cast<const short2 *>(iq)[i];
    return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
  }
  else return reinterpret_cast<const float2 *>(iq)[i];
}

__device__ __forceinline__ int mod_fs(long long x)
{
  int r = (int)(x % FS);
  return r < 0 ? r + FS : r;
}

// e^{j 2 pi idx / FS}: the reference's oscillator table entry (sample_reader.cpp:44-50), evaluated in double.
__device__ __forceinline__ float2 osc(int idx)
{
  double s, c;
  sincospi(2.0 * (double)idx / (double)FS, &s, &c);
  return make_float2((float)c, (float)s);
}

// Loads the 2048 samples starting at `start` in the FFT's register layout (v[n1] = x[128 n1 + tid]) and mixes them
// with the integer-Hz oscillator. `ph` is the oscillator phase BEFORE sample `start`. step[] is a 16-entry
// shared scratch (phasor of -f*128*n1).
template <int FMT>
__device__ __forceinline__ void load_symbol(float2 (&v)[16], const void * __restrict__ iq, long long n_total, long long start, int f, int ph,
                                            float2 * step, int tid)
{
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const long long i = start + 128 * n1 + tid;
    v[n1] = (i >= 0 && i < n_total) ? load_sample<FMT>(iq, i) : make_float2(0.0f, 0.0f);
  }
  if (f == 0 && ph == 0) return; // osc[0] = 1: the reference multiplies by exactly (1, 0)
  if (tid < 16) step[tid] = osc(mod_fs(-(long long)f * 128 * tid));
  __syncthreads();
  const float2 s0 = osc(mod_fs((long long)ph - (long long)f * (tid + 1)));
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++) v[n1] = cmul(v[n1], n1 == 0 ? s0 : cmul(s0, step[n1]));
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ FFT of whole frames
// Work item = (frame, row): row 0 = symbol 0, rows 1..75 data symbols, row 76 the null symbol.
// Output: X[slot][row][k] = spectrum at the bin of nominal carrier k (frequency de-interleaving fused into the store).
template <int FMT>
__global__ void __launch_bounds__(FFT_THREADS, 4) k_fft_frames(const FrameDesc * __restrict__ frames, int n_items, const RecInput * __restrict__ recs,
                                                               const float2 * __restrict__ w2048, const int16_t * __restrict__ bin_of_k,
                                                               float2 * __restrict__ X)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  __shared__ float2 step[16];
  __shared__ int16_t bins[K_CARR];
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int k = tid; k < K_CARR; k += FFT_THREADS) bins[k] = bin_of_k[k];
  __syncthreads();

  for (int item = blockIdx.x; item < n_items; item += gridDim.x)
  {
    const int fi = item / X_ROWS, row = item - fi * X_ROWS;
    const FrameDesc fd = frames[fi];
    if (row >= 1 && row <= 75 && row > fd.n_syms) continue;
    if (row == 76 && fd.n_syms < 75) continue;
    const RecInput rin = recs[fd.rec];
    long long start;
    int f, ph;
    if (row == 0) { start = fd.sym0; f = fd.f_sym0; ph = mod_fs((long long)fd.ph_eval - (long long)f * (fd.sym0 - fd.eval)); }
    else if (row <= 75)
    {
      start = fd.sym0 + T_U + (long long)(row - 1) * T_S + T_G;
      f = fd.f_data;
      ph = mod_fs((long long)fd.ph_data - (long long)f * ((long long)(row - 1) * T_S + T_G));
    }
    else { start = fd.sym0 + T_U + 75LL * T_S + T_G; f = fd.f_null; ph = mod_fs((long long)fd.ph_null - (long long)f * T_G); }

    float2 v[16];
    load_symbol<FMT>(v, rin.iq, rin.n, start, f, ph, step, tid);
    fft2048_to_smem(v, tw, smem, tid);
    float2 * out = X + ((size_t)fd.xslot * X_ROWS + row) * K_CARR;
#pragma unroll
    for (int i = 0; i < K_CARR / FFT_THREADS; i++)
    {
      const int k = tid + FFT_THREADS * i;
      out[k] = smem[bins[k]];
    }
    __syncthreads();
  }
}

// Natural-order batch transform (stage tap dabstar_fft2048). sign > 0: conj in, conj out.
__global__ void __launch_bounds__(FFT_THREADS, 4) k_fft_batch(const float2 * __restrict__ in, float2 * __restrict__ out, int n, int sign,
                                                              const float2 * __restrict__ w2048)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int item = blockIdx.x; item < n; item += gridDim.x)
  {
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++)
    {
      v[n1] = in[(size_t)item * T_U + 128 * n1 + tid];
      if (sign > 0) v[n1].y = -v[n1].y;
    }
    fft2048_to_smem(v, tw, smem, tid);
    for (int i = tid; i < T_U; i += FFT_THREADS)
    {
      float2 r = smem[i];
      if (sign > 0) r.y = -r.y;
      out[(size_t)item * T_U + i] = r;
    }
    __syncthreads();
  }
}

// Natural-order spectra [frame][77][2048] -> nominal-carrier order [frame][77][1536] (stage tap for the demapper).
__global__ void k_reorder_frames(const float2 * __restrict__ fft_nat, int n_rows, const int16_t * __restrict__ bin_of_k, float2 * __restrict__ X)
{
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  for (int k = threadIdx.x; k < K_CARR; k += blockDim.x) X[(size_t)row * K_CARR + k] = fft_nat[(size_t)row * T_U + bin_of_k[k]];
}

// ------------------------------------------------------------------------------------------------ block reductions
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ PRS correlation (S2)
// samples already in registers (FFT layout). Returns the start index to thread 0's caller through *result.
__device__ void prs_correlate_cta(float2 (&v)[16], const FftTwiddles & tw, const float2 * __restrict__ prs, float threshold, int strongest,
                                  float2 * smem, float * red, int * result, int tid)
{
  fft2048_to_smem(v, tw, smem, tid);
  // |IFFT(X conj(P))| = |FFT(conj(X) P)|
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const int i = 128 * n1 + tid;
    const float2 x = smem[i], p = prs[i];
    v[n1] = make_float2(x.x * p.x + x.y * p.y, x.x * p.y - x.y * p.x); // conj(x) * p
  }
  __syncthreads();
  fft2048_to_smem(v, tw, smem, tid);
  float * mag = reinterpret_cast<float *>(smem); // overwrite in place: mag[i] occupies the first half of smem[i]'s slot range
  float part = 0.0f;
  float m[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const float2 c = smem[128 * n1 + tid];
    m[n1] = sqrtf(c.x * c.x + c.y * c.y);
    part += m[n1];
  }
  __syncthreads();
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++) mag[128 * n1 + tid] = m[n1];
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  if (tid == 0) *result = 0x7fffffff;
  __syncthreads();
  const float mean = (red[0] + red[1] + red[2] + red[3]) / (float)T_U;
  if (mean == 0.0f) { if (tid == 0) *result = -1; __syncthreads(); return; }
  constexpr int lo = T_G - 250, hi = T_G + 500;
  if (!strongest)
  {
    // first index whose level exceeds the threshold and is not topped within the next 9 samples
    // (phasereference.cpp:141-168); the reference's skip-ahead only affects later candidates.
    int best = 0x7fffffff;
    for (int i = lo + tid; i < hi; i += FFT_THREADS)
    {
      const float a = mag[i];
      if (!(a / mean > threshold)) continue;
      bool peak = true;
      for (int j = 1; j < 10 && i + j < hi; j++) if (mag[i + j] > a) { peak = false; break; }
      if (peak) { best = i; break; }
    }
    if (best != 0x7fffffff) atomicMin(result, best);
    __syncthreads();
    if (tid == 0 && *result == 0x7fffffff) *result = -1;
  }
  else if (tid == 0)
  {
    int max_i = -1;
    float max_l = -1000.0f;
    for (int i = lo; i < hi; i++)
    {
      if (!(mag[i] / mean > threshold)) continue;
      bool peak = true;
      for (int j = 1; j < 10 && i + j < hi; j++) if (mag[i + j] > mag[i]) { peak = false; break; }
      if (!peak) continue;
      if (mag[i] > max_l) { max_l = mag[i]; max_i = i; }
      i += 10;
    }
    *result = (max_l / mean < threshold) ? -1 : max_i;
  }
  __syncthreads();
}

template <int FMT>
__global__ void __launch_bounds__(FFT_THREADS, 4) k_prs_corr(const FrameDesc * __restrict__ frames, int n_frames, const RecInput * __restrict__ recs,
                                                             const float2 * __restrict__ w2048, const float2 * __restrict__ prs,
                                                             float thr_first, float thr_rest, const uint8_t * __restrict__ first_flags,
                                                             int strongest, int * __restrict__ start_index)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  __shared__ float2 step[16];
  __shared__ float red[4];
  __shared__ int result;
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int fi = blockIdx.x; fi < n_frames; fi += gridDim.x)
  {
    const FrameDesc fd = frames[fi];
    const RecInput rin = recs[fd.rec];
    float2 v[16];
    load_symbol<FMT>(v, rin.iq, rin.n, fd.eval, fd.f_sym0, fd.ph_eval, step, tid);
    const float thr = (first_flags != nullptr && first_flags[fi]) ? thr_first : thr_rest;
    prs_correlate_cta(v, tw, prs, thr, strongest, smem, red, &result, tid);
    if (tid == 0) start_index[fi] = result;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(FFT_THREADS, 4) k_prs_corr_raw(const float2 * __restrict__ samples, int n, const float2 * __restrict__ w2048,
                                                                 const float2 * __restrict__ prs, float threshold, int strongest,
                                                                 int * __restrict__ start_index)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  __shared__ float red[4];
  __shared__ int result;
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int it = blockIdx.x; it < n; it += gridDim.x)
  {
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = samples[(size_t)it * T_U + 128 * n1 + tid];
    prs_correlate_cta(v, tw, prs, threshold, strongest, smem, red, &result, tid);
    if (tid == 0) start_index[it] = result;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ coarse AFC (S3)
// smem holds the natural-order spectrum of symbol 0 on entry.
__device__ void coarse_afc_cta(const FftTwiddles & tw, const float2 * __restrict__ ref_arg_conj, float2 * smem, float * magw, int * result, int tid)
{
  float2 v[16];
  // d[i] = conj(X[i]) X[i+1], d[2047] = 0; IFFT(d) = conj(FFT(conj(d)))
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const int i = 128 * n1 + tid;
    if (i < T_U - 1)
    {
      const float2 a = smem[i], b = smem[i + 1];
      const float2 d = make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
      v[n1] = make_float2(d.x, -d.y);
    }
    else v[n1] = make_float2(0.0f, 0.0f);
  }
  __syncthreads();
  fft2048_to_smem(v, tw, smem, tid);
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const int i = 128 * n1 + tid;
    const float2 c = make_float2(smem[i].x, -smem[i].y);
    v[n1] = cmul(c, ref_arg_conj[i]);
  }
  __syncthreads();
  fft2048_to_smem(v, tw, smem, tid);
  // peak over bins -70..70 (phasereference.cpp:248-279)
  for (int i = tid; i < 143; i += FFT_THREADS)
  {
    const float2 c = smem[(T_U + i - 71) & (T_U - 1)];
    magw[i] = sqrtf(c.x * c.x + c.y * c.y); // magw[i] = |bin i-71|, i = 0..142 (one guard bin each side)
  }
  __syncthreads();
  if (tid == 0)
  {
    int index = 100000;
    float mx = 0.0f, avg = 0.0f;
    for (int i = -70; i <= 70; i++)
    {
      const float val = magw[i + 71];
      if (val > mx) { mx = val; index = i; }
      avg += val;
    }
    avg /= 141.0f;
    if (mx < avg * 5.0f) *result = 100000;
    else
    {
      const float p0 = magw[index + 70], p1 = magw[index + 71], p2 = magw[index + 72];
      const float offset = (float)index + (p2 - p0) / (p0 + p1 + p2);
      *result = (int)(offset * 1000.0f);
    }
  }
  __syncthreads();
}

template <int FMT>
__global__ void __launch_bounds__(FFT_THREADS, 4) k_coarse_afc(const FrameDesc * __restrict__ frames, int n_frames, const RecInput * __restrict__ recs,
                                                               const float2 * __restrict__ w2048, const float2 * __restrict__ ref_arg_conj,
                                                               int * __restrict__ offset_hz)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  __shared__ float2 step[16];
  __shared__ float magw[144];
  __shared__ int result;
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int fi = blockIdx.x; fi < n_frames; fi += gridDim.x)
  {
    const FrameDesc fd = frames[fi];
    const RecInput rin = recs[fd.rec];
    float2 v[16];
    const int ph = mod_fs((long long)fd.ph_eval - (long long)fd.f_sym0 * (fd.sym0 - fd.eval));
    load_symbol<FMT>(v, rin.iq, rin.n, fd.sym0, fd.f_sym0, ph, step, tid);
    fft2048_to_smem(v, tw, smem, tid);
    coarse_afc_cta(tw, ref_arg_conj, smem, magw, &result, tid);
    if (tid == 0) offset_hz[fi] = result;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(FFT_THREADS, 4) k_coarse_afc_raw(const float2 * __restrict__ fft_nat, int n, const float2 * __restrict__ w2048,
                                                                   const float2 * __restrict__ ref_arg_conj, int * __restrict__ offset_hz)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  __shared__ float magw[144];
  __shared__ int result;
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  for (int it = blockIdx.x; it < n; it += gridDim.x)
  {
    for (int i = tid; i < T_U; i += FFT_THREADS) smem[i] = fft_nat[(size_t)it * T_U + i];
    __syncthreads();
    coarse_afc_cta(tw, ref_arg_conj, smem, magw, &result, tid);
    if (tid == 0) offset_hz[it] = result;
    __syncthreads();
  }
}

// ref_arg_conj = conj(IFFT(relative phase of the PRS)) (phasereference.cpp:58-66)
__global__ void __launch_bounds__(FFT_THREADS) k_init_ref_arg(const float2 * __restrict__ w2048, const float2 * __restrict__ prs, float2 * __restrict__ ref_arg_conj)
{
  __shared__ float2 smem[FFT_SMEM_F2];
  const int tid = threadIdx.x;
  FftTwiddles tw;
  fft_load_twiddles(tw, w2048, tid);
  float2 v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++)
  {
    const int i = 128 * n1 + tid;
    if (i < T_U - 1)
    {
      const float2 a = prs[i], b = prs[i + 1];
      v[n1] = make_float2(a.x * b.x + a.y * b.y, -(a.x * b.y - a.y * b.x)); // conj(conj(a) b)
    }
    else v[n1] = make_float2(0.0f, 0.0f);
  }
  fft2048_to_smem(v, tw, smem, tid);
  // IFFT(d) = conj(smem); ref_arg_conj = conj(IFFT(d)) = smem
  for (int i = tid; i < T_U; i += FFT_THREADS) ref_arg_conj[i] = smem[i];
}

// ------------------------------------------------------------------------------------------------ CP correlation (F2)
// cp[frame] = sum over symbols 1..75 and the 504 prefix samples of x[i+2048] conj(x[i]) on the RAW samples. The
// derotated sum the reference forms differs by the constant factor e^{-j 2 pi f / 1000} (f = integer Hz), which the
// host control loop applies.
template <int FMT>
__global__ void __launch_bounds__(256) k_cp_corr(const FrameDesc * __restrict__ frames, int n_frames, const RecInput * __restrict__ recs, float2 * __restrict__ cp)
{
  __shared__ float2 red[8];
  for (int fi = blockIdx.x; fi < n_frames; fi += gridDim.x)
  {
    const FrameDesc fd = frames[fi];
    const RecInput rin = recs[fd.rec];
    const long long base = fd.sym0 + T_U;
    float2 acc = make_float2(0.0f, 0.0f);
    const int total = fd.n_syms * T_G;
    for (int q = threadIdx.x; q < total; q += 256)
    {
      const int sym = q / T_G, i = q - sym * T_G;
      const long long p = base + (long long)sym * T_S + i;
      const float2 a = load_sample<FMT>(rin.iq, p + T_U), b = load_sample<FMT>(rin.iq, p);
      acc.x += a.x * b.x + a.y * b.y;
      acc.y += a.y * b.x - a.x * b.y;
    }
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      float2 s = red[0];
      for (int w = 1; w < 8; w++) { s.x += red[w].x; s.y += red[w].y; }
      cp[fi] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ DQPSK demapper (D1-D3)
// One CTA per recording walks its frames and symbols in order (the per-carrier IIR chains are sequential in time but
// independent across carriers). Thread t owns nominal carriers 2t and 2t+1 and keeps their state in registers.
constexpr int DEMAP_THREADS = K_CARR / 2;

__device__ __forceinline__ float first_quadrant(float ph)
{
  if (ph < 0.0f) ph += PI_F;            // common/glob_defs.h:173-182; fmod(x, pi/2) on [0, pi] is exact subtraction
  if (ph >= PI_2_F) ph -= PI_2_F;
  if (ph >= PI_2_F) ph -= PI_2_F;
  return ph;
}

// (i16)(float) as the reference's x86-64 build evaluates it (ofdm_decoder.cpp:254-255): cvttss2si to 32 bits
// (0x80000000 when out of range or NaN), then the low 16 bits. Out-of-range values only occur in the start-up
// transient of the SOFTDEC2 weighting; the Viterbi clamps to +-127 afterwards anyway.
__device__ __forceinline__ short to_i16(float v)
{
  const int r = (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
  return (short)(r & 0xffff);
}

struct CarrierState
{
  float integ, stddev, mean_pow, mean_sigma, null_pow;
};

// decode_symbol for one carrier (ofdm_decoder.cpp:166-294). Returns r1 (soft value before the symbol-wide scale).
template <int SOFT>
__device__ __forceinline__ float2 demap_carrier(CarrierState & st, float2 x, float2 ref, float clock_term)
{
  constexpr float ALPHA = 0.005f;
  const float ref_abs = sqrtf(ref.x * ref.x + ref.y * ref.y);
  float2 raw = cmul_conj(x, ref);
  raw.x /= ref_abs;
  raw.y /= ref_abs;
  const float perr = clock_term + st.integ;
  const float a = -perr, a2 = a * a;
  const float2 rot = make_float2(0.9994032382965087890625f + a2 * (a2 * 3.679168224334716796875e-2f + -0.495580852031707763671875f),
                                 a * (a2 * -0.16034401953220367431640625f + 0.99903142452239990234375f));
  const float2 z = cmul(raw, rot);
  const float ph = first_quadrant(atan2f(z.y, z.x));
  st.integ += 0.2f * ALPHA * (ph - PI_4_F);
  st.integ = fminf(fmaxf(st.integ, -20.0f * RAD_PER_DEG_F), 20.0f * RAD_PER_DEG_F);
  const float dv = ph - PI_4_F;
  st.stddev += ALPHA * (dv * dv - st.stddev);
  const float pw = z.x * z.x + z.y * z.y;
  st.mean_pow += ALPHA * (pw - st.mean_pow);
  const float lvl = sqrtf(st.mean_pow);
  const float axis = lvl * 0.70710678118654752440f;
  const float dr = fabsf(z.x) - axis, di = fabsf(z.y) - axis;
  st.mean_sigma += ALPHA * (dr * dr + di * di - st.mean_sigma);
  float sig = st.mean_pow - st.null_pow;
  if (sig <= 0.0f) sig = 0.1f;
  float w1;
  if (SOFT == 2) w1 = ref_abs;
  else if (SOFT == 1)
  {
    w1 = ref_abs / st.mean_sigma;
    w1 /= st.null_pow / sig + 0.7f;
  }
  else
  {
    const float zabs = sqrtf(pw);
    w1 = sqrtf(zabs * ref_abs) * lvl;
    w1 /= st.null_pow / sig + 0.7f;
    w1 /= st.mean_sigma * zabs;
  }
  return make_float2(z.x * w1, z.y * w1);
}

template <int SOFT>
__global__ void __launch_bounds__(DEMAP_THREADS, 1) k_demap(const DemapWork * __restrict__ work, const FrameDesc * __restrict__ frames,
                                                            const uint8_t * __restrict__ null_is_tii, const float2 * __restrict__ X,
                                                            const int16_t * __restrict__ rel_of_k, OfdmStateDev * __restrict__ states,
                                                            int16_t * __restrict__ soft)
{
  __shared__ float red[2][DEMAP_THREADS / 32];
  const DemapWork wk = work[blockIdx.x];
  const int tid = threadIdx.x, k0 = 2 * tid;
  OfdmStateDev & sd = states[wk.state];
  CarrierState s0, s1;
  if (wk.reset)
  {
    s0 = CarrierState{ 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
    s1 = s0;
  }
  else
  {
    s0 = CarrierState{ sd.integ[k0], sd.stddev[k0], sd.mean_pow[k0], sd.mean_sigma[k0], sd.null_pow[k0] };
    s1 = CarrierState{ sd.integ[k0 + 1], sd.stddev[k0 + 1], sd.mean_pow[k0 + 1], sd.mean_sigma[k0 + 1], sd.null_pow[k0 + 1] };
  }
  float mean_value = sd.mean_value; // not touched by reset() (ofdm_decoder.cpp:90-101)
  const float g0 = (float)(K_CARR / 2 - rel_of_k[k0]) / (float)(K_CARR / 2);
  const float g1 = (float)(K_CARR / 2 - rel_of_k[k0 + 1]) / (float)(K_CARR / 2);
  constexpr float W2 = SOFT == 0 ? -100.0f : -140.0f;
  int buf = 0;

  for (int fi = 0; fi < wk.n_frames; fi++)
  {
    const FrameDesc fd = frames[wk.desc_first + fi];
    const float4 * row = reinterpret_cast<const float4 *>(X + (size_t)fd.xslot * X_ROWS * K_CARR) + tid;
    int16_t * out = soft + (size_t)fd.slot * FRAME_SOFT;
    const float ce = fd.clock_err / 1024.0f * PI_F;
    const float c0 = ce * g0, c1 = ce * g1;
    float4 ref = row[0];                                   // store_reference_symbol_0
    float4 cur = row[K_CARR / 2];                          // symbol 1 (row stride = 1536 float2 = 768 float4)
    for (int sym = 1; sym <= fd.n_syms; sym++)
    {
      float4 nxt = cur;
      if (sym < 76) nxt = row[(size_t)(sym + 1) * (K_CARR / 2)]; // prefetch symbol sym+1 (row 76 = null symbol)
      const float2 r0 = demap_carrier<SOFT>(s0, make_float2(cur.x, cur.y), make_float2(ref.x, ref.y), c0);
      const float2 r1 = demap_carrier<SOFT>(s1, make_float2(cur.z, cur.w), make_float2(ref.z, ref.w), c1);
      const float w2 = W2 / mean_value;
      short2 re = make_short2(to_i16(r0.x * w2), to_i16(r1.x * w2));
      short2 im = make_short2(to_i16(r0.y * w2), to_i16(r1.y * w2));
      int16_t * o = out + (size_t)(sym - 1) * SYM_BITS;
      reinterpret_cast<short2 *>(o)[tid] = re;
      reinterpret_cast<short2 *>(o + K_CARR)[tid] = im;
      // mMeanValue = sum |r1| / K for the NEXT symbol (ofdm_decoder.cpp:256,294)
      float part = sqrtf(r0.x * r0.x + r0.y * r0.y) + sqrtf(r1.x * r1.x + r1.y * r1.y);
      part = warp_sum(part);
      if ((tid & 31) == 0) red[buf][tid >> 5] = part;
      __syncthreads();
      float tot = 0.0f;
#pragma unroll
      for (int w = 0; w < DEMAP_THREADS / 32; w++) tot += red[buf][w];
      mean_value = tot / (float)K_CARR;
      buf ^= 1;
      ref = cur;
      cur = nxt;
    }
    // null symbol: store_null_symbol_without_tii (ofdm_decoder.cpp:114-130); `cur` holds row 76 when the frame is complete
    if (fd.n_syms == 75 && !(null_is_tii != nullptr && null_is_tii[wk.desc_first + fi]))
    {
      constexpr float MIN_POW = (1.0f / 32767.0f) * (1.0f / 32767.0f);
      const float p0 = cur.x * cur.x + cur.y * cur.y + MIN_POW, p1 = cur.z * cur.z + cur.w * cur.w + MIN_POW;
      s0.null_pow += 0.05f * (p0 - s0.null_pow);
      s1.null_pow += 0.05f * (p1 - s1.null_pow);
    }
  }
  sd.integ[k0] = s0.integ; sd.integ[k0 + 1] = s1.integ;
  sd.stddev[k0] = s0.stddev; sd.stddev[k0 + 1] = s1.stddev;
  sd.mean_pow[k0] = s0.mean_pow; sd.mean_pow[k0 + 1] = s1.mean_pow;
  sd.mean_sigma[k0] = s0.mean_sigma; sd.mean_sigma[k0 + 1] = s1.mean_sigma;
  sd.null_pow[k0] = s0.null_pow; sd.null_pow[k0 + 1] = s1.null_pow;
  if (tid == 0) sd.mean_value = mean_value;
}

// ------------------------------------------------------------------------------------------------ time sync (S1)
// TimeSyncer::read_samples_until_end_of_level_drop (timesyncer.cpp:40-90) on top of SampleReader's level IIR
// (sample_reader.cpp:236, alpha = 1e-5). One CTA per recording; the stream is scanned in blocks of 1024 samples.
constexpr int DIP_THREADS = 256;
constexpr int DIP_BLOCK = 1024;
constexpr float LEVEL_ALPHA = 0.00001f;

template <int FMT> __device__ __forceinline__ float sample_abs(const void * iq, long long i)
{
  const float2 v = load_sample<FMT>(iq, i);
  return sqrtf(v.x * v.x + v.y * v.y);
}

template <int FMT>
__global__ void __launch_bounds__(DIP_THREADS) k_dip_search(const DipWork * __restrict__ work, const RecInput * __restrict__ recs, DipResult * __restrict__ out)
{
  __shared__ float mag[DIP_BLOCK + 64];
  __shared__ float lvl[DIP_BLOCK];
  __shared__ float2 scan_ab[DIP_THREADS];
  __shared__ float redf[DIP_THREADS / 32];
  __shared__ int found;
  const DipWork wk = work[blockIdx.x];
  const RecInput rin = recs[wk.rec];
  const int tid = threadIdx.x;
  const long long p0 = wk.pos;
  const float loga = log1pf(-LEVEL_ALPHA);

  // sLevel at p0: IIR over everything read so far (initial value 0.1, sample_reader.h:91); older than 2^21 samples is < 1e-9.
  const long long hist = p0 < (1LL << 21) ? p0 : (1LL << 21);
  float acc = 0.0f;
  for (long long i = tid; i < hist; i += DIP_THREADS) acc += sample_abs<FMT>(rin.iq, p0 - 1 - i) * expf((float)i * loga);
  acc = warp_sum(acc);
  if ((tid & 31) == 0) redf[tid >> 5] = acc;
  __syncthreads();
  float s_level = 0.0f;
  for (int w = 0; w < DIP_THREADS / 32; w++) s_level += redf[w];
  s_level = LEVEL_ALPHA * s_level + (hist == p0 ? 0.1f * expf((float)p0 * loga) : 0.0f);
  __syncthreads();

  // phase 0: searching the dip (first check after 50 samples), phase 1: searching its end
  int phase = 0;
  long long dip_i = -1;
  int status = -1;
  long long end_pos = p0;
  const float a1 = 1.0f - LEVEL_ALPHA;
  for (long long b0 = 0; status < 0; b0 += DIP_BLOCK)
  {
    if (p0 + b0 + DIP_BLOCK > rin.n) { status = 3; end_pos = rin.n; break; }
    // |x| for this block plus the 49 preceding samples (window history)
    for (int i = tid; i < DIP_BLOCK + 49; i += DIP_THREADS)
    {
      const long long g = b0 + i - 49;
      mag[i] = g >= 0 ? sample_abs<FMT>(rin.iq, p0 + g) : 0.0f;
    }
    __syncthreads();
    // level IIR s_i = a s_{i-1} + alpha |x_i| as a scan of affine maps, 4 samples per thread
    float A = 1.0f, B = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) { const float m = mag[49 + 4 * tid + q]; B = a1 * B + LEVEL_ALPHA * m; A *= a1; }
    scan_ab[tid] = make_float2(A, B);
    __syncthreads();
    if (tid == 0)
    {
      float s = s_level;
      for (int t = 0; t < DIP_THREADS; t++) { const float2 ab = scan_ab[t]; scan_ab[t].x = s; s = ab.x * s + ab.y; }
      scan_ab[0].y = s; // level after the block
    }
    __syncthreads();
    {
      float s = scan_ab[tid].x;
#pragma unroll
      for (int q = 0; q < 4; q++) { s = a1 * s + LEVEL_ALPHA * mag[49 + 4 * tid + q]; lvl[4 * tid + q] = s; }
    }
    if (tid == 0) found = 0x7fffffff;
    __syncthreads();
    const float s_after = scan_ab[0].y;
    // evaluate both phases inside this block (the dip end may follow in the same block)
    int from = 0;
    while (true)
    {
      int best = 0x7fffffff;
      for (int q = 0; q < 4; q++)
      {
        const int i = 4 * tid + q;
        const long long gi = b0 + i;
        if (i < from) continue;
        if (phase == 0 && gi < 49) continue;
        float c = 0.0f;
        for (int j = 0; j < 50; j++) c += mag[i + j]; // mag[i..i+49] = samples gi-49..gi
        const float mean = c / 50.0f;
        const bool hit = phase == 0 ? !(mean > 0.55f * lvl[i]) : !(mean < 0.75f * lvl[i]);
        if (hit) { best = i; break; }
      }
      if (best != 0x7fffffff) atomicMin(&found, best);
      __syncthreads();
      const int f = found;
      __syncthreads();
      if (tid == 0) found = 0x7fffffff;
      __syncthreads();
      // limits: dip within T_F loop reads (timesyncer.cpp:66-69), end within T_N + 70 (timesyncer.cpp:82-85)
      if (phase == 0)
      {
        const long long lim = 49 + (long long)T_F; // last index whose condition is still evaluated
        const long long gi = f == 0x7fffffff ? (long long)1 << 60 : b0 + f;
        if (gi <= lim) { phase = 1; dip_i = gi; from = f; continue; }
        if (b0 + DIP_BLOCK - 1 >= lim + 1) { status = 1; end_pos = p0 + lim + 2; }
        break;
      }
      else
      {
        const long long lim = dip_i + T_N + 70;
        const long long gi = f == 0x7fffffff ? (long long)1 << 60 : b0 + f;
        if (gi <= lim) { status = 0; end_pos = p0 + gi + 1; }
        else if (b0 + DIP_BLOCK - 1 >= lim + 1) { status = 2; end_pos = p0 + lim + 2; }
        break;
      }
    }
    s_level = s_after;
    __syncthreads();
  }
  if (tid == 0)
  {
    if (end_pos > rin.n) { end_pos = rin.n; status = 3; }
    out[blockIdx.x].pos = end_pos;
    out[blockIdx.x].status = status;
    out[blockIdx.x].s_level = s_level;
  }
}

template <typename F> cudaError_t dispatch_fmt(int fmt, F && f)
{
  switch (fmt)
  {
  case FMT_CF32: f(std::integral_constant<int, FMT_CF32>()); break;
  case FMT_U8: f(std::integral_constant<int, FMT_U8>()); break;
  case FMT_I16: f(std::integral_constant<int, FMT_I16>()); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

inline int fft_grid(int n_items) { return n_items < N_SM * 4 ? n_items : N_SM * 4; }
} // namespace

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_init_ref_arg(cudaStream_t s, const DeviceTables & t, unsigned long long * lc)
{
  k_init_ref_arg<<<1, FFT_THREADS, 0, s>>>(t.w2048, t.prs, t.ref_arg_conj);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_fft_batch(cudaStream_t s, const DeviceTables & t, const float2 * in, float2 * out, int n, int sign, unsigned long long * lc)
{
  if (n <= 0) return cudaSuccess;
  k_fft_batch<<<fft_grid(n), FFT_THREADS, 0, s>>>(in, out, n, sign, t.w2048);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_fft_frames(cudaStream_t s, const DeviceTables & t, const FrameDesc * frames, int n_frames, const RecInput * recs, int fmt,
                              float2 * X, unsigned long long * lc)
{
  if (n_frames <= 0) return cudaSuccess;
  const int n_items = n_frames * X_ROWS;
  if (lc) (*lc)++;
  return dispatch_fmt(fmt, [&](auto F) { k_fft_frames<decltype(F)::value><<<fft_grid(n_items), FFT_THREADS, 0, s>>>(frames, n_items, recs, t.w2048, t.bin_of_k, X); });
}

cudaError_t launch_reorder_frames(cudaStream_t s, const DeviceTables & t, const float2 * fft_nat, int n_frames, float2 * X, unsigned long long * lc)
{
  if (n_frames <= 0) return cudaSuccess;
  k_reorder_frames<<<n_frames * X_ROWS, 256, 0, s>>>(fft_nat, n_frames * X_ROWS, t.bin_of_k, X);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_demap(cudaStream_t s, const DeviceTables & t, const DemapWork * work, int n_work, const FrameDesc * frames,
                         const uint8_t * null_is_tii, const float2 * X, OfdmStateDev * states, int soft_bit_type, int16_t * soft,
                         unsigned long long * lc)
{
  if (n_work <= 0) return cudaSuccess;
  if (lc) (*lc)++;
  switch (soft_bit_type)
  {
  case 0: k_demap<0><<<n_work, DEMAP_THREADS, 0, s>>>(work, frames, null_is_tii, X, t.rel_of_k, states, soft); break;
  case 1: k_demap<1><<<n_work, DEMAP_THREADS, 0, s>>>(work, frames, null_is_tii, X, t.rel_of_k, states, soft); break;
  case 2: k_demap<2><<<n_work, DEMAP_THREADS, 0, s>>>(work, frames, null_is_tii, X, t.rel_of_k, states, soft); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_cp_corr(cudaStream_t s, const FrameDesc * frames, int n_frames, const RecInput * recs, int fmt, float2 * cp, unsigned long long * lc)
{
  if (n_frames <= 0) return cudaSuccess;
  if (lc) (*lc)++;
  const int grid = n_frames < N_SM * 8 ? n_frames : N_SM * 8;
  return dispatch_fmt(fmt, [&](auto F) { k_cp_corr<decltype(F)::value><<<grid, 256, 0, s>>>(frames, n_frames, recs, cp); });
}

cudaError_t launch_prs_corr(cudaStream_t s, const DeviceTables & t, const FrameDesc * frames, int n_frames, const RecInput * recs, int fmt,
                            float threshold_first, float threshold_rest, const uint8_t * first_flags, int strongest, int * start_index, unsigned long long * lc)
{
  if (n_frames <= 0) return cudaSuccess;
  if (lc) (*lc)++;
  return dispatch_fmt(fmt, [&](auto F) {
    k_prs_corr<decltype(F)::value><<<fft_grid(n_frames), FFT_THREADS, 0, s>>>(frames, n_frames, recs, t.w2048, t.prs, threshold_first, threshold_rest, first_flags, strongest, start_index);
  });
}

cudaError_t launch_prs_corr_raw(cudaStream_t s, const DeviceTables & t, const float2 * samples, int n, float threshold, int strongest,
                                int * start_index, unsigned long long * lc)
{
  if (n <= 0) return cudaSuccess;
  k_prs_corr_raw<<<fft_grid(n), FFT_THREADS, 0, s>>>(samples, n, t.w2048, t.prs, threshold, strongest, start_index);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_coarse_afc(cudaStream_t s, const DeviceTables & t, const FrameDesc * frames, int n_frames, const RecInput * recs, int fmt,
                              int * offset_hz, unsigned long long * lc)
{
  if (n_frames <= 0) return cudaSuccess;
  if (lc) (*lc)++;
  return dispatch_fmt(fmt, [&](auto F) {
    k_coarse_afc<decltype(F)::value><<<fft_grid(n_frames), FFT_THREADS, 0, s>>>(frames, n_frames, recs, t.w2048, t.ref_arg_conj, offset_hz);
  });
}

cudaError_t launch_coarse_afc_raw(cudaStream_t s, const DeviceTables & t, const float2 * fft_nat, int n, int * offset_hz, unsigned long long * lc)
{
  if (n <= 0) return cudaSuccess;
  k_coarse_afc_raw<<<fft_grid(n), FFT_THREADS, 0, s>>>(fft_nat, n, t.w2048, t.ref_arg_conj, offset_hz);
  if (lc) (*lc)++;
  return cudaGetLastError();
}

cudaError_t launch_dip_search(cudaStream_t s, const DipWork * work, int n, const RecInput * recs, int fmt, DipResult * out, unsigned long long * lc)
{
  if (n <= 0) return cudaSuccess;
  if (lc) (*lc)++;
  return dispatch_fmt(fmt, [&](auto F) { k_dip_search<decltype(F)::value><<<n, DIP_THREADS, 0, s>>>(work, recs, out); });
}
} // namespace dab
