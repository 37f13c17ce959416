// tables.cu — constant tables of the Mode-I path, built on the host at context creation and kept in HBM.
// They restate ETSI EN 300 401 tables the reference also carries: frequency interleaver
// (ofdm/freq_interleaver.cpp:40-76), phase reference symbol (ofdm/phasetable.cpp:35-135), puncturing vectors
// (protection/protTables.cpp:36-62), UEP profiles (protection/uep_protection.cpp:52-132), energy-dispersal PRBS
// (decoder/fic_decoder.cpp:59-73).
#include "tables.h"
#include "fft2048.cuh"
#include <vector>

#include <cmath>
#include <cstring>

namespace dab
{
static const uint32_t PI_MASK[24] = {
  0x11111113u, 0x11131113u, 0x11131313u, 0x13131313u, 0x13131333u, 0x13331333u, 0x13333333u, 0x33333333u,
  0x33333337u, 0x33373337u, 0x33373737u, 0x37373737u, 0x37373777u, 0x37773777u, 0x37777777u, 0x77777777u,
  0x7777777fu, 0x777f777fu, 0x777f7f7fu, 0x7f7f7f7fu, 0x7f7f7fffu, 0x7fff7fffu, 0x7fffffffu, 0xffffffffu };

struct UepRow { int16_t bitrate, level, L[4], PI[4]; };
static const UepRow UEP_TABLE[] = {
  {32,5, {3,4,17,0}, {5,3,2,0}}, {32,4, {3,3,18,0}, {11,6,5,0}}, {32,3, {3,4,14,3}, {15,9,6,8}},
  {32,2, {3,4,14,3}, {22,13,8,13}}, {32,1, {3,5,13,3}, {24,17,12,17}}, {48,5, {4,3,26,3}, {5,4,2,3}},
  {48,4, {3,4,26,3}, {9,6,4,6}}, {48,3, {3,4,26,3}, {15,10,6,9}}, {48,2, {3,4,26,3}, {24,14,8,15}},
  {48,1, {3,5,25,3}, {24,18,13,18}}, {56,5, {6,10,23,3}, {5,4,2,3}}, {56,4, {6,10,23,3}, {9,6,4,5}},
  {56,3, {6,12,21,3}, {16,7,6,9}}, {56,2, {6,10,23,3}, {23,13,8,13}}, {64,5, {6,9,31,2}, {5,3,2,3}},
  {64,4, {6,9,33,0}, {11,6,5,0}}, {64,3, {6,12,27,3}, {16,8,6,9}}, {64,2, {6,10,29,3}, {23,13,8,13}},
  {64,1, {6,11,28,3}, {24,18,12,18}}, {80,5, {6,10,41,3}, {6,3,2,3}}, {80,4, {6,10,41,3}, {11,6,5,6}},
  {80,3, {6,11,40,3}, {16,8,6,7}}, {80,2, {6,10,41,3}, {23,13,8,13}}, {80,1, {6,10,41,3}, {24,7,12,18}},
  {96,5, {7,9,53,3}, {5,4,2,4}}, {96,4, {7,10,52,3}, {9,6,4,6}}, {96,3, {6,12,51,3}, {16,9,6,10}},
  {96,2, {6,10,53,3}, {22,12,9,12}}, {96,1, {6,13,50,3}, {24,18,13,19}}, {112,5, {14,17,50,3}, {5,4,2,5}},
  {112,4, {11,21,49,3}, {9,6,4,8}}, {112,3, {11,23,47,3}, {16,8,6,9}}, {112,2, {11,21,49,3}, {23,12,9,14}},
  {128,5, {12,19,62,3}, {5,3,2,4}}, {128,4, {11,21,61,3}, {11,6,5,7}}, {128,3, {11,22,60,3}, {16,9,6,10}},
  {128,2, {11,21,61,3}, {22,12,9,14}}, {128,1, {11,20,62,3}, {24,17,13,19}}, {160,5, {11,19,87,3}, {5,4,2,4}},
  {160,4, {11,23,83,3}, {11,6,5,9}}, {160,3, {11,24,82,3}, {16,8,6,11}}, {160,2, {11,21,85,3}, {22,11,9,13}},
  {160,1, {11,22,84,3}, {24,18,12,19}}, {192,5, {11,20,110,3}, {6,4,2,5}}, {192,4, {11,22,108,3}, {10,6,4,9}},
  {192,3, {11,24,106,3}, {16,10,6,11}}, {192,2, {11,20,110,3}, {22,13,9,13}}, {192,1, {11,21,109,3}, {24,20,13,24}},
  {224,5, {12,22,131,3}, {8,6,2,6}}, {224,4, {12,26,127,3}, {12,8,4,11}}, {224,3, {11,20,134,3}, {16,10,7,9}},
  {224,2, {11,22,132,3}, {24,16,10,15}}, {224,1, {11,24,130,3}, {24,20,12,20}}, {256,5, {11,24,154,3}, {6,5,2,5}},
  {256,4, {11,24,154,3}, {12,9,5,10}}, {256,3, {11,27,151,3}, {16,10,7,10}}, {256,2, {11,22,156,3}, {24,14,10,13}},
  {256,1, {11,26,152,3}, {24,19,14,18}}, {320,5, {11,26,200,3}, {8,5,2,6}}, {320,4, {11,25,201,3}, {13,9,5,10}},
  {320,2, {11,26,200,3}, {24,17,9,17}}, {384,5, {11,27,247,3}, {8,6,2,7}}, {384,3, {11,24,250,3}, {16,9,7,10}},
  {384,1, {12,28,245,3}, {24,20,14,23}} };
static const int N_UEP = (int)(sizeof(UEP_TABLE) / sizeof(UEP_TABLE[0]));

static const uint8_t PRS_ROW[48] = { 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3,
                                     0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1 };
static const uint8_t PRS_N[48] = { 1,2,0,1, 3,2,2,3, 2,1,2,3, 1,2,3,3, 2,2,2,1, 1,3,1,2,
                                   3,1,1,1, 2,2,1,0, 2,2,3,3, 0,2,1,3, 3,3,3,0, 3,0,1,1 };
static const uint8_t PRS_H[4][32] = {
  { 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1, 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1 },
  { 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0, 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0 },
  { 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3, 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3 },
  { 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2, 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2 } };

void host_freq_interleaver(int16_t * bin_signed)
{
  int n = 0, v = 0;
  for (int i = 0; i < T_U; i++)
  {
    if (i > 0) v = (13 * v + 511) % T_U;
    if (v == T_U / 2 || v < 256 || v > 256 + K_CARR) continue;
    bin_signed[n++] = (int16_t)(v - T_U / 2);
  }
}

// Staging layout of the FFT kernel's epilogue. After the last stage thread tid holds the bins fft_out_index(tid, h, j2) as its
// values i = 8 h + j2; the frequency de-interleaver (a pseudo-random permutation, freq_interleaver.cpp:37-60) sends them to
// nominal carriers; the demapper's rows want carriers 2p, 2p + 1 next to each other. The values pass through shared memory once:
// written with STS.64 (16 per thread: value i of the 32 lanes of a warp at a time), read with LDS.64 (carrier 2p + b of 32
// consecutive p at a time). A 64-bit access is served per HALF warp, 16 lanes on 16 pairs of banks, so an instruction is
// conflict free when the 16 slots of each half warp differ modulo 16. Every carrier sits in exactly one write group (warp,
// half, i) and one read group (iteration, b, warp, half), groups hold at most 16 carriers: the carriers are the edges of a
// bipartite multigraph of maximum degree 16, which has a proper edge colouring with 16 colours (Koenig); colour = slot modulo
// 16, slot = 16 * write group + colour. Written with the pseudo-random slots of the straightforward layout (slot = carrier) a
// scatter store took 6.5 wavefronts instead of 2 (ncu, profiles/r2_m_*), a third of the kernel's shared-memory traffic.
void host_fft_epilogue_layout(const int16_t * bin_of_k, uint16_t * slot_w, uint16_t * slot_r)
{
  constexpr int NW = 128, NR = 96, D = 16;
  std::vector<int> carrier_of_bin(T_U, -1);
  for (int k = 0; k < K_CARR; k++) carrier_of_bin[bin_of_k[k]] = k;
  // edges
  std::vector<int> wg(K_CARR, -1), rg(K_CARR, -1);
  for (int i = 0; i < 16; i++)
    for (int tid = 0; tid < FFT_THREADS; tid++)
    {
      const int c = carrier_of_bin[fft_out_index(tid, i >> 3, i & 7)];
      if (c >= 0) wg[c] = (i * 4 + (tid >> 5)) * 2 + ((tid >> 4) & 1);
    }
  for (int c = 0; c < K_CARR; c++)
  {
    const int p = c >> 1, b = c & 1, it = p / FFT_THREADS, tid = p % FFT_THREADS;
    rg[c] = ((it * 2 + b) * 4 + (tid >> 5)) * 2 + ((tid >> 4) & 1);
  }
  // bipartite edge colouring with D colours: at[side][vertex][colour] = edge using that colour at the vertex, or -1
  std::vector<int> at_w(NW * D, -1), at_r(NR * D, -1), colour(K_CARR, -1);
  for (int e = 0; e < K_CARR; e++)
  {
    const int u = wg[e], v = rg[e];
    int a = 0, b = 0;
    while (at_w[u * D + a] >= 0) a++; // free at u
    while (at_r[v * D + b] >= 0) b++; // free at v
    if (a != b)
    {
      // colour a is taken at v: flip the a / b alternating path that starts there (it cannot reach u, where a is free and the
      // path would have to arrive through an a edge)
      std::vector<int> path;
      int x = v, side = 1, want = a; // at a right vertex looking for its `a` edge
      while (true)
      {
        const int f = side ? at_r[x * D + want] : at_w[x * D + want];
        if (f < 0) break;
        path.push_back(f);
        x = side ? wg[f] : rg[f];
        side ^= 1;
        want = want == a ? b : a;
      }
      for (int f : path) { at_w[wg[f] * D + colour[f]] = -1; at_r[rg[f] * D + colour[f]] = -1; }
      for (int f : path) { colour[f] = colour[f] == a ? b : a; }
      for (int f : path) { at_w[wg[f] * D + colour[f]] = f; at_r[rg[f] * D + colour[f]] = f; }
    }
    colour[e] = a;
    at_w[u * D + a] = e;
    at_r[v * D + a] = e;
  }
  for (int c = 0; c < K_CARR; c++) slot_r[c] = (uint16_t)(wg[c] * D + colour[c]);
  // the writers: carriers go to their slots, values of bins without a carrier to the colours their write group has left
  for (int i = 0; i < 16; i++)
    for (int tid = 0; tid < FFT_THREADS; tid++)
    {
      const int g = (i * 4 + (tid >> 5)) * 2 + ((tid >> 4) & 1);
      const int c = carrier_of_bin[fft_out_index(tid, i >> 3, i & 7)];
      if (c >= 0) { slot_w[i * FFT_THREADS + tid] = slot_r[c]; continue; }
      int a = 0;
      while (at_w[g * D + a] >= 0) a++;
      at_w[g * D + a] = K_CARR; // taken by a value nobody reads
      slot_w[i * FFT_THREADS + tid] = (uint16_t)(g * D + a);
    }
}

void host_phase_table(float2 * prs)
{
  const float pi_2 = (float)(M_PI / 2.0);
  memset(prs, 0, sizeof(float2) * T_U);
  for (int k = -K_CARR / 2; k <= K_CARR / 2; k++)
  {
    if (k == 0) continue;
    const int blk = k < 0 ? (k + 768) / 32 : 24 + (k - 1) / 32;
    const int kmin = k < 0 ? -768 + 32 * blk : 1 + 32 * (blk - 24);
    const float phi = pi_2 * (float)(PRS_H[PRS_ROW[blk]][k - kmin] + PRS_N[blk]);
    prs[k < 0 ? k + T_U : k] = make_float2(cosf(phi), sinf(phi)); // cmplx_from_phase (glob_defs.h:139-143)
  }
}

void host_prbs(uint8_t * out, int n)
{
  unsigned reg = 0x1ff;
  for (int i = 0; i < n; i++)
  {
    const unsigned b = ((reg >> 8) ^ (reg >> 4)) & 1u;
    reg = ((reg << 1) | b) & 0x1ff;
    out[i] = (uint8_t)b;
  }
}

void host_w2048(float2 * w)
{
  for (int m = 0; m < T_U; m++) w[m] = make_float2((float)cos(2.0 * M_PI * m / T_U), (float)-sin(2.0 * M_PI * m / T_U));
}

static void add_segment(VitProfile & p, int blocks, int pi)
{
  if (blocks <= 0) return;
  const int s = p.n_seg++;
  const uint32_t m = PI_MASK[pi - 1];
  p.seg_mask[s] = m;
  p.seg_kept[s] = p.n_kept;
  p.seg_pos[s + 1] = p.seg_pos[s] + 128 * blocks;
  p.n_kept += 4 * blocks * __builtin_popcount(m);
}

static void add_tail(VitProfile & p)
{
  const int s = p.n_seg++;
  p.seg_mask[s] = PI_MASK[7] & 0x00ffffffu; // PI_X = first 24 entries of PI_8
  p.seg_kept[s] = p.n_kept;
  p.seg_pos[s + 1] = p.seg_pos[s] + 24;
  p.n_kept += __builtin_popcount(p.seg_mask[s]);
}

VitProfile make_fic_profile()
{
  VitProfile p;
  memset(&p, 0, sizeof(p));
  p.n_bits = FIC_OUT;
  add_segment(p, 21, 16);
  add_segment(p, 3, 15);
  add_tail(p);
  return p;
}

VitProfile make_identity_profile(int n_bits)
{
  VitProfile p;
  memset(&p, 0, sizeof(p));
  p.n_bits = n_bits;
  p.n_seg = 1;
  p.seg_mask[0] = 0xffffffffu;
  p.seg_pos[1] = 4 * (n_bits + 6);
  p.n_kept = 4 * (n_bits + 6);
  return p;
}

bool make_msc_profile(int short_form, int bit_rate, int prot_level, VitProfile & p)
{
  memset(&p, 0, sizeof(p));
  if (bit_rate <= 0 || bit_rate > 1024) return false;
  p.n_bits = 24 * bit_rate;
  if (short_form)
  {
    int idx = -1;
    for (int i = 0; i < N_UEP; i++) if (UEP_TABLE[i].bitrate == bit_rate && UEP_TABLE[i].level == prot_level) { idx = i; break; }
    if (idx < 0) return false; // the reference silently substitutes row 1 (uep_protection.cpp:163-167); refuse instead
    int blocks = 0;
    for (int s = 0; s < 4; s++) blocks += UEP_TABLE[idx].L[s];
    if (blocks != 6 * bit_rate / 8) return false;
    for (int s = 0; s < 4; s++) if (UEP_TABLE[idx].PI[s] > 0) add_segment(p, UEP_TABLE[idx].L[s], UEP_TABLE[idx].PI[s]);
  }
  else
  {
    if (prot_level < 0 || prot_level > 7) return false;
    const int level = prot_level & 3;
    int L1, L2, P1, P2;
    if ((prot_level & 4) == 0)
    {
      if (bit_rate % 8) return false;
      const int n = bit_rate / 8;
      switch (level)
      {
      case 0: L1 = 6 * n - 3; L2 = 3; P1 = 24; P2 = 23; break;
      case 1: if (n == 1) { L1 = 5; L2 = 1; P1 = 13; P2 = 12; } else { L1 = 2 * n - 3; L2 = 4 * n + 3; P1 = 14; P2 = 13; } break;
      case 2: L1 = 6 * n - 3; L2 = 3; P1 = 8; P2 = 7; break;
      default: L1 = 4 * n - 3; L2 = 2 * n + 3; P1 = 3; P2 = 2; break;
      }
    }
    else
    {
      if (bit_rate % 32) return false;
      static const int8_t pb[4][2] = { {10, 9}, {6, 5}, {4, 3}, {2, 1} };
      L1 = 24 * (bit_rate / 32) - 3; L2 = 3; P1 = pb[level][0]; P2 = pb[level][1];
    }
    add_segment(p, L1, P1);
    add_segment(p, L2, P2);
  }
  add_tail(p);
  return p.seg_pos[p.n_seg] == 4 * (p.n_bits + 6);
}

int profile_addresses(const VitProfile & p, int32_t * addr, int cap)
{
  int n = 0;
  for (int pos = 0; pos < 4 * (p.n_bits + 6); pos++)
  {
    const int idx = vit_src_index(p, pos);
    if (idx < 0) continue;
    if (idx < cap) addr[idx] = pos;
    n++;
  }
  return n;
}
} // namespace dab
