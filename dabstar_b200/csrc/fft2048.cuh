// fft2048.cuh — 2048-point complex FFT executed by one CTA of 128 threads, 16 points per thread.
//
// Replaces the reference's FFTW3f plan (main/dab_processor.cpp:63; ofdm/phasereference.cpp:51-52):
// unnormalised forward DFT, X[k] = sum_n x[n] e^{-j 2 pi k n / 2048}. The backward transform the
// reference also uses is obtained by conjugating input and output.
//
// Decomposition 2048 = 16 x 16 x 8 (Cooley-Tukey, decimation in time over the input index):
//   n = 128 n1 + n2,           k = k1 + 16 k2            (n1,k1 < 16; n2,k2 < 128)
//   n2 = 8 m1 + m2,            k2 = j1 + 16 j2           (m1,j1 < 16; m2,j2 < 8)
//   stage 1: thread t = n2      : 16-pt DFT over n1, times W2048^(t k1)        -> SA[k1][t]
//   stage 2: thread u = 8 k1+m2 : 16-pt DFT over m1, times W128^(m2 j1)        -> SB[k1][j1][m2]
//   stage 3: thread owns pairs p = 16 k1 + j1 (p = u and u + 128): 8-pt DFT over m2 -> X[k1 + 16 j1 + 256 j2]
// All butterflies are in registers; two shared-memory exchanges; the per-thread twiddles are loaded
// once and stay in registers while a persistent CTA walks over many symbols.
//
// Every phase is a __host__ __device__ function of (thread id, registers, shared array) so that
// tests/test_fft_emulation.py can run the exact index arithmetic on the CPU (no GPU in the build box).
#pragma once
#include "common.cuh"

namespace dab
{
constexpr int FFT_THREADS = 128;
constexpr int FFT_SA_STRIDE = 136;  // float2 per k1 row, 136*2 words = 16 mod 32: half-warp rows land on disjoint banks
constexpr int FFT_SB_STRIDE = 152;  // float2 per k1 row (16 j1 x 9 + 8): same property for the stage-2 stores
constexpr int FFT_SB_J1 = 9;        // 8 m2 + 1 pad: stage-3 LDS.64 of consecutive pairs hit distinct banks
constexpr int FFT_SMEM_F2 = 16 * FFT_SB_STRIDE; // 2432 float2 = 19 456 B (>= 2048 for the natural-order staging)

constexpr int FFT_TW2_F2 = 16 * 8;  // second-stage twiddle table W128^(m2*j1), [j1][m2], shared by the CTA

struct FftTwiddles
{
  float2 w1[16];      // W2048^(t*k1), t = tid: per thread, in registers
  const float2 * w2;  // shared-memory table [j1][m2] = W128^(m2*j1) = W2048^(16*m2*j1)
};

// w2048[m] = e^{-j 2 pi m / 2048}, computed on the host in double (tables.cu). tw2 must hold FFT_TW2_F2 entries;
// every thread of the CTA calls this (the caller synchronises before the first transform).
__host__ __device__ inline void fft_load_twiddles(FftTwiddles & tw, const float2 * w2048, float2 * tw2, int tid)
{
#pragma unroll
  for (int i = 0; i < 16; i++) tw.w1[i] = w2048[(tid * i) & 2047];
  tw2[tid] = w2048[(16 * (tid & 7) * (tid >> 3)) & 2047]; // tid = 8*j1 + m2
  tw.w2 = tw2;
}

// Natural-order staging index: one pad element per 16 so that the stage-3 scatter (stride 16) and the strided
// gathers of the callers spread over the banks.
__host__ __device__ inline int fft_nat(int n) { return n + (n >> 4); }
constexpr int FFT_NAT_F2 = 2048 + 128;

__host__ __device__ inline void dft4(float2 & a, float2 & b, float2 & c, float2 & d)
{
  const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  a = cadd(t0, t2);
  c = csub(t0, t2);
  b = csub_j(t1, t3); // t1 - j t3
  d = cadd_j(t1, t3); // t1 + j t3
}

// In-place 16-point forward DFT, natural order in and out.
__host__ __device__ inline void dft16(float2 (&v)[16])
{
  constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, R = 0.70710678118654752440f;
  // n = 4a + b: DFT over a for each b -> v[4c + b] holds T[b][c]
#pragma unroll
  for (int b = 0; b < 4; b++) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // twiddles W16^(b c)
  v[4 + 1] = cmul_const(v[4 + 1], C1, -S1);   // b=1,c=1 : m=1
  v[8 + 1] = cmul_const(v[8 + 1], R, -R);     // b=1,c=2 : m=2
  v[12 + 1] = cmul_const(v[12 + 1], S1, -C1); // b=1,c=3 : m=3
  v[4 + 2] = cmul_const(v[4 + 2], R, -R);     // b=2,c=1 : m=2
  v[8 + 2] = cmul_mj(v[8 + 2]);               // b=2,c=2 : m=4 -> -j
  v[12 + 2] = cmul_const(v[12 + 2], -R, -R);  // b=2,c=3 : m=6
  v[4 + 3] = cmul_const(v[4 + 3], S1, -C1);   // b=3,c=1 : m=3
  v[8 + 3] = cmul_const(v[8 + 3], -R, -R);    // b=3,c=2 : m=6
  v[12 + 3] = cmul_const(v[12 + 3], -C1, S1); // b=3,c=3 : m=9
  // DFT over b for each c -> X[c + 4d] in slot v[4c + d]
#pragma unroll
  for (int c = 0; c < 4; c++) dft4(v[4 * c + 0], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  // slot 4c + d holds X[c + 4d]: transpose the 4x4 to natural order
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int d = c + 1; d < 4; d++)
    {
      const float2 t = v[4 * c + d];
      v[4 * c + d] = v[4 * d + c];
      v[4 * d + c] = t;
    }
}

// In-place 8-point forward DFT on v[o..o+7], natural order in and out.
__host__ __device__ inline void dft8(float2 * v)
{
  constexpr float R = 0.70710678118654752440f;
  // n = 2a + b: DFT over a (4 points) for b = 0,1
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  o1 = cmul_const(o1, R, -R);
  o3 = cmul_const(o3, -R, -R);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = csub_j(e2, o2); v[6] = cadd_j(e2, o2); // o2 * (-j) folded into the add
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// Phase 1: v[n1] = x[128 n1 + tid] on entry.
__host__ __device__ inline void fft_stage1(float2 (&v)[16], const FftTwiddles & tw, float2 * smem, int tid)
{
  dft16(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; k1++)
  {
    const float2 a = k1 == 0 ? v[0] : cmul(v[k1], tw.w1[k1]);
    smem[k1 * FFT_SA_STRIDE + tid] = a;
  }
}

// Phase 2a (after a barrier): gather the 16 inputs of this thread's second-level DFT.
__host__ __device__ inline void fft_stage2_read(float2 (&v)[16], const float2 * smem, int tid)
{
  const int k1 = tid >> 3, m2 = tid & 7;
#pragma unroll
  for (int m1 = 0; m1 < 16; m1++) v[m1] = smem[k1 * FFT_SA_STRIDE + 8 * m1 + m2];
}

// Phase 2b (after a barrier: every thread has read SA): DFT, twiddle, store SB.
__host__ __device__ inline void fft_stage2(float2 (&v)[16], const FftTwiddles & tw, float2 * smem, int tid)
{
  const int k1 = tid >> 3, m2 = tid & 7;
  dft16(v);
#pragma unroll
  for (int j1 = 0; j1 < 16; j1++)
  {
    const float2 b = j1 == 0 ? v[0] : cmul(v[j1], tw.w2[8 * j1 + m2]);
    smem[k1 * FFT_SB_STRIDE + j1 * FFT_SB_J1 + m2] = b;
  }
}

// Phase 3 (after a barrier): two 8-point DFTs. On return v[8 h + j2] = X[fft_out_index(tid, h, j2)].
__host__ __device__ inline void fft_stage3(float2 (&v)[16], const float2 * smem, int tid)
{
#pragma unroll
  for (int h = 0; h < 2; h++)
  {
    const int p = tid + 128 * h, k1 = p >> 4, j1 = p & 15;
#pragma unroll
    for (int m2 = 0; m2 < 8; m2++) v[8 * h + m2] = smem[k1 * FFT_SB_STRIDE + j1 * FFT_SB_J1 + m2];
    dft8(&v[8 * h]);
  }
}

__host__ __device__ inline int fft_out_index(int tid, int h, int j2)
{
  const int p = tid + 128 * h;
  return (p >> 4) + 16 * (p & 15) + 256 * j2;
}

#ifdef __CUDACC__
// Whole transform for a CTA of FFT_THREADS threads. v[n1] = x[128 n1 + tid] on entry; on return the
// natural-order spectrum is in smem[fft_nat(0..2047)] (all threads synchronised).
__device__ inline void fft2048_to_smem(float2 (&v)[16], const FftTwiddles & tw, float2 * smem, int tid)
{
  fft_stage1(v, tw, smem, tid);
  __syncthreads();
  fft_stage2_read(v, smem, tid);
  __syncthreads();
  fft_stage2(v, tw, smem, tid);
  __syncthreads();
  fft_stage3(v, smem, tid);
  __syncthreads();
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int j2 = 0; j2 < 8; j2++) smem[fft_nat(fft_out_index(tid, h, j2))] = v[8 * h + j2];
  __syncthreads();
}
#endif
} // namespace dab
