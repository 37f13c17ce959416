// viterbi_tpc.cuh — K=7 rate-1/4 Viterbi decoder, one THREAD per terminated code word ("tpc").
//
// Same function as viterbi.cuh (ViterbiSpiral::deconvolve, support/viterbi_spiral/viterbi_spiral.cpp:95-126,
// viterbi_scalar.h:9-94: start metrics 0 / 1000, metric = sum_j (BT_j ^ s_j), path 0 wins ties, chain back from
// state 0), organised for the integer pipes of sm_100a instead of for one warp:
//
//   * the 64 path metrics of a code word live in 32 registers as unsigned 16-bit PAIRS. One VIMNMX.U16x2 (the packed
//     integer add-and-min of sm_100) advances one path and selects the survivors of two states, one 32-bit add advances
//     the other path, and the two decisions are the sign bits of the packed difference m0 - m1 + 0x7fff: one PRMT in
//     sign-replicate mode gathers the four decisions of a packed butterfly, one LOP3 files them in the decision word.
//     The adds are written as two-input adds wherever possible so that ptxas can spread them over the ALU and the FMA
//     pipe (IADD3 / IMAD.IADD): the ALU pipe alone (min, PRMT, LOP3) bounds the kernel. A warp therefore advances 32
//     code words with ~200 instructions per trellis step, where the warp-per-code-word kernel needs ~60 instructions
//     per step for ONE code word.
//   * Exactness: the reference's 32-bit metrics never differ by more than 6*1020+1000 = 7120 between states. The
//     packed metrics are renormalised (minimum subtracted) every TPC_RENORM steps, so they stay below
//     7120 + (TPC_RENORM + 5) * 1020 < 65536: no half ever carries into its neighbour, unsigned comparison is exact and
//     the decisions are those of the reference.
//   * Register pairing. Butterfly a (old states a, a+32 -> new states 2a, 2a+1) is paired with butterfly a^m in the
//     two halves of a register; the results then pair states that differ by 2m, which is exactly the pairing the next
//     step needs with mask 2m. The mask walks 1,2,4,8,16; after the fifth step the pairs are (x, x+32) and one PRMT per
//     register restores mask 1. The five step types are five template instances, everything is indexed at compile
//     time, nothing is spilled.
//   * Branch metrics: generators 109 (twice), 79, 83 give 8 distinct metrics per step; butterflies a and a^m differ by
//     a fixed pattern flip per step type, so 8 packed constants per step serve all 16 packed butterflies
//     (1020 - metric is the constant of the complementary pattern).
//   * The decision word of step t goes to global memory, [step][code word] so that a warp writes 256 contiguous
//     bytes; chain back reads it the same way. The decision of new state y sits at bit tpc_decision_bit(K, y) (the
//     order in which the PRMT/LOP3 pairs file them), a 5 x 64 byte table for the chain back.
//
// Every function is __host__ __device__ so tests/native/host_logic.cu runs the identical code on the CPU.
#pragma once
#include "viterbi.cuh"

namespace dab
{
constexpr int TPC_RENORM = 40; // steps between renormalisations (multiple of 5)

__host__ __device__ constexpr int tpc_par(unsigned x) { return (int)(((x ^ (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4) ^ (x >> 5) ^ (x >> 6) ^ (x >> 7)) & 1u)); }
// metric pattern of butterfly a: bit 0 = generator 109 (symbols 0 and 3), bit 1 = generator 79, bit 2 = generator 83
__host__ __device__ constexpr int tpc_pat(int a) { return tpc_par((2u * a) & 109u) | (tpc_par((2u * a) & 79u) << 1) | (tpc_par((2u * a) & 83u) << 2); }
// index a with bit k removed
__host__ __device__ constexpr int tpc_compress(int a, int k) { return (a & ((1 << k) - 1)) | ((a >> (k + 1)) << k); }

__host__ __device__ __forceinline__ unsigned tpc_minu(unsigned a, unsigned b)
{
#ifdef __CUDA_ARCH__
  return __vminu2(a, b); // VIMNMX.U16x2
#else
  const unsigned al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
  return (bl < al ? bl : al) | ((bh < ah ? bh : ah) << 16);
#endif
}

// min(a + b, c) per 16-bit half: VIADDMNMX.U16x2, one ALU-pipe instruction for the add and the select
__host__ __device__ __forceinline__ unsigned tpc_addmin(unsigned a, unsigned b, unsigned c)
{
#ifdef __CUDA_ARCH__
  return __vminu2(__vadd2(a, b), c);
#else
  return tpc_minu(a + b, c); // no half of a + b carries (see "Exactness" above)
#endif
}

// Bytes [t0.b1, t1.b1, t0.b3, t1.b3] with the most significant bit of each replicated over the byte (0x00 / 0xFF).
__host__ __device__ __forceinline__ unsigned tpc_signs(unsigned t0, unsigned t1)
{
#ifdef __CUDA_ARCH__
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0xFBD9;" : "=r"(r) : "r"(t0), "r"(t1));
  return r;
#else
  return ((t0 >> 15) & 1u ? 0xffu : 0u) | ((t1 >> 15) & 1u ? 0xff00u : 0u) | ((t0 >> 31) & 1u ? 0xff0000u : 0u) | ((t1 >> 31) & 1u ? 0xff000000u : 0u);
#endif
}

// (low half of x, low half of y) and (high half of x, high half of y): PRMT on the device
__host__ __device__ __forceinline__ unsigned tpc_lo_lo(unsigned x, unsigned y)
{
#ifdef __CUDA_ARCH__
  return __byte_perm(x, y, 0x5410);
#else
  return (x & 0xffffu) | (y << 16);
#endif
}
__host__ __device__ __forceinline__ unsigned tpc_hi_hi(unsigned x, unsigned y)
{
#ifdef __CUDA_ARCH__
  return __byte_perm(x, y, 0x7632);
#else
  return (x >> 16) | (y & 0xffff0000u);
#endif
}
// byte j of x
template <int J> __host__ __device__ __forceinline__ unsigned tpc_byte(unsigned x)
{
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0u, 0x4440 + J);
#else
  return (x >> (8 * J)) & 255u;
#endif
}
__host__ __device__ __forceinline__ unsigned tpc_lo_hi(unsigned x, unsigned y) { return (x & 0xffffu) | (y & 0xffff0000u); }

// Bit of the decision word of a type-K step that holds the decision of new state y (see TpcButterfly).
__host__ __device__ constexpr int tpc_decision_bit(int K, int y)
{
  const int c = y >> 1, half = (c >> K) & 1, a = c & ~(1 << K), i = tpc_compress(a, K);
  return 32 * (i >> 3) + 8 * (2 * half + (y & 1)) + (i & 7);
}

// Start metrics in the layout of step type 0 (viterbi_spiral.cpp:98-101): state 0 -> 0, every other state -> 1000.
__host__ __device__ __forceinline__ void tpc_init(unsigned (&S)[32])
{
#pragma unroll
  for (int i = 0; i < 32; i++) S[i] = 1000u | (1000u << 16);
  S[0] = 0u | (1000u << 16); // (old[0], old[1])
}

// One trellis step of type K (pair mask 1 << K). Layout of S: for butterfly a with bit K clear, i = tpc_compress(a, K):
//   S[i] = old[a] | old[a ^ m] << 16,   S[16 + i] = old[a + 32] | old[(a ^ m) + 32] << 16.
// T receives the new metrics in the layout of type K + 1; for K = 4 in the layout T[x] = new[x] | new[x + 32] << 16.
// sym: the four clamped symbols of the step, byte j = symbol j. dlo/dhi: decision word (tpc_decision_bit).
// Butterflies with i < TPC_DFORM take their two decision words from D = L - H (three two-input adds, which ptxas may
// place on either integer pipe), the others from one three-input add each (ALU pipe only): see tpc_step.
constexpr int TPC_DFORM = 16;

template <int K, int A>
struct TpcButterfly
{
  static __host__ __device__ __forceinline__ void run(const unsigned (&S)[32], unsigned (&T)[32], const unsigned (&PB)[8], const unsigned (&KB)[8], unsigned & lo,
                                                      unsigned & hi)
  {
    if constexpr (((A >> K) & 1) == 0)
    {
      constexpr int i = tpc_compress(A, K), p = tpc_pat(A);
      const unsigned L = S[i], H = S[16 + i];
      // low halves: butterfly A (new states 2A, 2A+1); high halves: butterfly A ^ m.
      // m0 = L + PB[p], m1 = H + PB[p ^ 7] -> new state 2A; m2 = L + PB[p ^ 7], m3 = H + PB[p] -> new state 2A + 1
      const unsigned m1 = H + PB[p ^ 7], m3 = H + PB[p];
      const unsigned n0 = tpc_addmin(L, PB[p], m1), n1 = tpc_addmin(L, PB[p ^ 7], m3);
      // decision = 1 when the path from old state +32 is strictly better (viterbi_scalar.h: d = (m0 - m1) > 0):
      // per half m0 - m1 + 0x7fff has bit 15 set exactly then (|m0 - m1| <= 7120 + 1020, so the halves of the final
      // value neither borrow nor carry; the 32-bit intermediates may, which cancels). With PB[p] + PB[p ^ 7] = 1020 per
      // half, m0 - m1 + 0x7fff = (L - H) + KB[p] and m2 - m3 + 0x7fff = (L - H) + KB[p ^ 7], KB[p] = 2 PB[p] + (0x7fff - 1020).
      unsigned t0, t1;
      if constexpr (i < TPC_DFORM)
      {
        const unsigned D = L - H;
        t0 = D + KB[p];
        t1 = D + KB[p ^ 7];
      }
      else
      {
        t0 = L + (PB[p] + 0x7fff7fffu) - m1;
        t1 = L + (PB[p ^ 7] + 0x7fff7fffu) - m3;
      }
      const unsigned x = tpc_signs(t0, t1) & (0x01010101u << (i & 7));
      if constexpr (i < 8) lo |= x; else hi |= x;
      if constexpr (K < 4)
      {
        if constexpr (A < 16) { T[tpc_compress(2 * A, K + 1)] = n0; T[tpc_compress(2 * A + 1, K + 1)] = n1; }
        else { T[16 + tpc_compress(2 * A - 32, K + 1)] = n0; T[16 + tpc_compress(2 * A - 31, K + 1)] = n1; }
      }
      else { T[2 * A] = n0; T[2 * A + 1] = n1; }
    }
    if constexpr (A + 1 < 32) TpcButterfly<K, A + 1>::run(S, T, PB, KB, lo, hi);
  }
};

template <int K>
__host__ __device__ __forceinline__ void tpc_step(const unsigned (&S)[32], unsigned (&T)[32], unsigned sym, unsigned & dlo, unsigned & dhi)
{
  // packed branch metrics: low half = metric of pattern p, high half = metric of pattern p ^ F (the partner butterfly).
  // Each packed term is one multiply-add of the symbol value x (maximum M): x * 0x00010001 = (x, x),
  // x * 0xfffeffff + (M, M) = (M - x, M - x), x * 0xffff0001 + (0, M) = (x, M - x), x * 0x0000ffff + (M, 0) = (M - x, x)
  // (low half first), which keeps them on the FMA pipe.
  constexpr int F = tpc_pat(1 << K);
  const unsigned s1 = tpc_byte<1>(sym), s2 = tpc_byte<2>(sym), sa = tpc_byte<0>(sym) + tpc_byte<3>(sym);
  const unsigned Ap[2] = { (F & 1) ? sa * 0xffff0001u + 0x01fe0000u : sa * 0x00010001u, (F & 1) ? sa * 0x0000ffffu + 0x000001feu : sa * 0xfffeffffu + 0x01fe01feu };
  const unsigned Bp[2] = { (F & 2) ? s1 * 0xffff0001u + 0x00ff0000u : s1 * 0x00010001u, (F & 2) ? s1 * 0x0000ffffu + 0x000000ffu : s1 * 0xfffeffffu + 0x00ff00ffu };
  const unsigned Cp[2] = { (F & 4) ? s2 * 0xffff0001u + 0x00ff0000u : s2 * 0x00010001u, (F & 4) ? s2 * 0x0000ffffu + 0x000000ffu : s2 * 0xfffeffffu + 0x00ff00ffu };
  unsigned PB[8], KB[8];
#pragma unroll
  for (int p = 0; p < 8; p++)
  {
    PB[p] = Ap[p & 1] + Bp[(p >> 1) & 1] + Cp[(p >> 2) & 1];
    KB[p] = PB[p] * 2u + 0x7c037c03u; // 0x7fff - 1020 per half
  }
  unsigned lo = 0, hi = 0;
  TpcButterfly<K, 0>::run(S, T, PB, KB, lo, hi);
  dlo = lo;
  dhi = hi;
}

// (x, x + 32) pairs -> type-0 layout; RENORM: the minimum metric is subtracted from every state.
template <bool RENORM>
__host__ __device__ __forceinline__ void tpc_repair(const unsigned (&R)[32], unsigned (&S)[32])
{
  unsigned sub = 0;
  if constexpr (RENORM)
  {
    unsigned m = R[0];
#pragma unroll
    for (int i = 1; i < 32; i++) m = tpc_minu(m, R[i]);
    const unsigned ml = m & 0xffffu, mh = m >> 16, mn = ml < mh ? ml : mh;
    sub = mn | (mn << 16);
  }
#pragma unroll
  for (int a = 0; a < 32; a += 2)
  {
    const unsigned x = RENORM ? R[a] - sub : R[a], y = RENORM ? R[a + 1] - sub : R[a + 1]; // no half borrows: every half >= mn
    S[a >> 1] = tpc_lo_lo(x, y);      // (s[a], s[a+1])
    S[16 + (a >> 1)] = tpc_hi_hi(x, y); // (s[a+32], s[a+33])
  }
}

// Five steps t0..t0+4 (types 0..4) and the re-pairing. sym[i]: symbols of step t0 + i. dec[i]: decision word of step t0 + i.
__host__ __device__ __forceinline__ void tpc_five_steps(unsigned (&S)[32], const unsigned (&sym)[5], unsigned long long (&dec)[5], bool renorm)
{
  unsigned T[32], lo, hi;
  tpc_step<0>(S, T, sym[0], lo, hi); dec[0] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  tpc_step<1>(T, S, sym[1], lo, hi); dec[1] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  tpc_step<2>(S, T, sym[2], lo, hi); dec[2] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  tpc_step<3>(T, S, sym[3], lo, hi); dec[3] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  tpc_step<4>(S, T, sym[4], lo, hi); dec[4] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  if (renorm) tpc_repair<true>(T, S); else tpc_repair<false>(T, S);
}

// Chain back one step: y = current state, w = decision word of the step that entered y, pos = tpc_decision_bit of
// (type of that step, y). Returns the decoded bit.
__host__ __device__ __forceinline__ unsigned tpc_traceback_step(unsigned long long w, unsigned pos, unsigned & y)
{
  const unsigned k = (unsigned)(w >> pos) & 1u;
  y = (y >> 1) | (k << 5);
  return k;
}

// The four clamped symbols of trellis step t of a code word (viterbi_scalar.h:34-40: `i16 tmp = in; tmp += 127` wraps
// in 16 bits before the clamp to [0, 255]); v[g] = soft input at Viterbi position 4 t + g, 0 where punctured.
__host__ __device__ __forceinline__ unsigned tpc_pack_syms(const int (&v)[4])
{
  unsigned r = 0;
#pragma unroll
  for (int g = 0; g < 4; g++)
  {
    int x = (int)(short)(v[g] + 127);
    x = x < 0 ? 0 : (x > 255 ? 255 : x);
    r |= (unsigned)x << (8 * g);
  }
  return r;
}
} // namespace dab
