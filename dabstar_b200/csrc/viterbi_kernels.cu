// viterbi_kernels.cu — warp-per-code-word Viterbi with fused depuncture / time de-interleave gather,
// energy dispersal, FIB CRC and BER re-encode. See viterbi.cuh for the algorithm and reference citations.
#include "viterbi.cuh"
#include "kernels.h"

namespace dab
{
namespace
{
constexpr unsigned FULL = 0xffffffffu;
constexpr int VIT_SMEM_PER_STEP = 14; // bytes of shared memory per trellis step and warp

__device__ __forceinline__ unsigned clamp_sym(int v)
{
  // viterbi_scalar.h:34-40: `i16 tmp = in; tmp += 127;` wraps in 16 bits before the clamp (inputs above 32640 become 0)
  v = (int)(short)(v + 127);
  return (unsigned)min(max(v, 0), 255);
}

__device__ __forceinline__ unsigned parity8(unsigned x)
{
  return __popc(x) & 1u;
}

// shared memory per warp: survivors u64[cap] | symbols u32[cap] | decoded bits u8[cap] | input signs, 1 bit per position (cap/2 bytes)
// fic_frames != nullptr: job j is FIC block (j & 3) of frame descriptor j >> 2 (no job list needed)
__global__ void __launch_bounds__(256) k_viterbi(const VitJob * __restrict__ jobs, const FrameDesc * __restrict__ fic_frames, int n_jobs, const VitProfile * __restrict__ profiles,
                                                 const int16_t * __restrict__ soft, uint8_t * __restrict__ out_bits,
                                                 const uint8_t * __restrict__ prbs, uint8_t * __restrict__ crc_ok,
                                                 int * __restrict__ ber, int cap)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  unsigned char * base = smem_raw + (size_t)warp * (size_t)cap * VIT_SMEM_PER_STEP;
  unsigned long long * surv = reinterpret_cast<unsigned long long *>(base);
  unsigned * syms = reinterpret_cast<unsigned *>(base + (size_t)cap * 8);
  unsigned char * syms8 = reinterpret_cast<unsigned char *>(syms);
  unsigned char * dbits = base + (size_t)cap * 12;
  unsigned * sgn = reinterpret_cast<unsigned *>(base + (size_t)cap * 13);

  const unsigned xmask = vit_branch_mask(lane);
  const unsigned sel = (lane & 1) ? 0x3244u : 0x1044u; // PRMT: chosen 16-bit half -> upper half, zeros below
  const int src_a = lane >> 1, src_b = 16 + (lane >> 1);

  for (int j = blockIdx.x * warps + warp; j < n_jobs; j += gridDim.x * warps)
  {
    VitJob job;
    if (fic_frames != nullptr)
    {
      const FrameDesc & fd = fic_frames[j >> 2];
      const int b = j & 3, n_fic = min(4, fd.n_syms * SYM_BITS / FIC_IN);
      if (b >= n_fic) continue;
      job.src = (long long)fd.slot * FRAME_SOFT + (long long)b * FIC_IN;
      job.out = (long long)fd.slot * (4 * FIC_OUT) + (long long)b * FIC_OUT;
      job.profile = 0;
      job.src_mode = VIT_SRC_LINEAR;
      job.flags = VIT_FLAG_PRBS | VIT_FLAG_FIC;
      job.cif_first = job.row_mask = job.frag_off = 0;
      job.aux = 4 * fd.slot + b;
    }
    else job = jobs[j];
    const VitProfile pr = profiles[job.profile];
    const int n_bits = pr.n_bits, steps = n_bits + 6;

    // ---- gather: depuncture (+ time de-interleave) and clamp to u8
    for (int p0 = 0; p0 < 4 * steps; p0 += 32)
    {
      const int p = p0 + lane;
      int v = 0;
      if (p < 4 * steps)
      {
        const int idx = vit_src_index(pr, p);
        if (idx >= 0)
        {
          if (job.src_mode == VIT_SRC_LINEAR) v = soft[job.src + idx];
          else
          {
            const int m = time_map(idx);
            if ((job.row_mask >> m) & 1) v = soft[job.src + cif_offset(job.cif_first + m) + job.frag_off + idx];
          }
        }
        syms8[p] = (unsigned char)clamp_sym(v);
      }
      const unsigned pos = __ballot_sync(FULL, v > 0); // sign of the raw input, for calculate_BER
      if (lane == 0) sgn[p0 >> 5] = pos;
    }
    __syncwarp();

    // ---- forward pass
    unsigned a = (lane == 0 ? 0u : 1000u) << 16; // old[i]   (start metrics 0 / 1000, viterbi_spiral.cpp:98-101)
    unsigned b = 1000u << 16;                    // old[i+32]
    for (int t = 0; t < steps; t++)
    {
      const unsigned metric = __dp4a(syms[t] ^ xmask, 0x01010101u, 0u);
      unsigned n0, n1;
      bool d0, d1;
      vit_acs(a, b, metric, n0, n1, d0, d1);
      const unsigned w0 = __ballot_sync(FULL, d0), w1 = __ballot_sync(FULL, d1);
      if (lane == 0) surv[t] = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
      const unsigned pk = __byte_perm(n0, n1, 0x7632); // low half = n0 >> 16, high half = n1 >> 16
      const unsigned va = __shfl_sync(FULL, pk, src_a), vb = __shfl_sync(FULL, pk, src_b);
      a = __byte_perm(va, 0u, sel);
      b = __byte_perm(vb, 0u, sel);
    }
    __syncwarp();

    // ---- chain back from state 0, 32 decisions per round
    unsigned r = 0;
    for (int top = steps - 1; top >= 6; top -= 32)
    {
      const int e = top - lane;
      const unsigned long long w = e >= 0 ? surv[e] : 0ull;
      unsigned mine = 0;
#pragma unroll 8
      for (int q = 0; q < 32; q++)
      {
        const unsigned long long wq = __shfl_sync(FULL, w, q);
        const unsigned k = vit_traceback_step(wq, r);
        if (q == lane) mine = k;
      }
      const int t = e - 6;
      if (t >= 0)
      {
        if (job.flags & VIT_FLAG_FIC) dbits[t] = (unsigned char)mine;
        out_bits[job.out + t] = (unsigned char)(mine ^ ((job.flags & VIT_FLAG_PRBS) ? prbs[t] : 0));
      }
    }

    if (job.flags & VIT_FLAG_FIC)
    {
      __syncwarp();
      // BER: re-encode and compare with the sign of the kept inputs (viterbi_spiral.cpp:128-164)
      int errors = 0;
      for (int i = lane; i < steps; i += 32)
      {
        unsigned sr = 0;
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
          const int ii = i - q;
          if (ii >= 0 && ii < n_bits) sr |= (unsigned)dbits[ii] << q;
        }
        const int poly[4] = { 109, 79, 83, 109 };
#pragma unroll
        for (int g = 0; g < 4; g++)
        {
          if (vit_src_index(pr, 4 * i + g) < 0) continue;
          const unsigned hard = (sgn[(4 * i + g) >> 5] >> ((4 * i + g) & 31)) & 1u;
          errors += hard != parity8(sr & (unsigned)poly[g]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) errors += __shfl_xor_sync(FULL, errors, o);
      if (lane == 0) { ber[2 * job.aux] = pr.n_kept; ber[2 * job.aux + 1] = errors; }
      // CRC of the three FIBs on the descrambled bits (backend/crc.cpp:98-132)
      if (lane < 3)
      {
        unsigned reg = 0xffff;
        const int o = lane * 256;
        for (int i = 0; i < 256; i++)
        {
          const unsigned bit = (unsigned)(dbits[o + i] ^ prbs[o + i]) ^ (i >= 240 ? 1u : 0u);
          const unsigned fb = ((reg >> 15) & 1u) ^ bit;
          reg = (reg << 1) & 0xffffu;
          if (fb) reg ^= 0x1021u;
        }
        crc_ok[3 * job.aux + lane] = reg == 0;
      }
    }
    __syncwarp();
  }
}
} // namespace

int viterbi_smem_bytes(int max_steps, int warps)
{
  const int cap = (max_steps + 15) & ~15;
  return cap * VIT_SMEM_PER_STEP * warps;
}

// Picks warps per CTA so the shared-memory footprint allows several CTAs per SM, launches the jobs.
cudaError_t launch_viterbi(cudaStream_t stream, const VitJob * jobs, const FrameDesc * fic_frames, int n_jobs, const VitProfile * profiles, int max_steps,
                           const int16_t * soft, uint8_t * out_bits, const uint8_t * prbs, uint8_t * crc_ok, int * ber,
                           unsigned long long * launch_counter)
{
  if (n_jobs <= 0) return cudaSuccess;
  const int cap = (max_steps + 15) & ~15;
  const int per_warp = cap * VIT_SMEM_PER_STEP;
  int warps = 8;
  while (warps > 1 && per_warp * warps > 56 * 1024) warps >>= 1; // <= 56 KB per CTA -> 4 CTAs/SM when possible
  const int smem = per_warp * warps;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set)
  {
    cudaError_t e = cudaFuncSetAttribute(k_viterbi, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int ctas_needed = (n_jobs + warps - 1) / warps;
  const int per_sm = max(1, min(16, (227 * 1024) / max(smem, 1)));
  const int grid = min(ctas_needed, N_SM * per_sm);
  k_viterbi<<<grid, warps * 32, smem, stream>>>(jobs, fic_frames, n_jobs, profiles, soft, out_bits, prbs, crc_ok, ber, cap);
  if (launch_counter) (*launch_counter)++;
  return cudaGetLastError();
}
} // namespace dab
