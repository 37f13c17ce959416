// Host-only checks of the product's table builders and of the warp Viterbi algorithm (emulated lanes).
// Built with nvcc as a host program (no GPU needed). Protocol on stdout, parsed by tests/test_native_host.py:
//   ADDR <short_form> <prot_level> <bit_rate> <count> <fnv1a of the address list>
//   VIT <frame_bits> <seed> <fnv1a of decoded bits>      for inputs read from stdin (binary int16)
#include "tables.h"
#include "viterbi.cuh"
#include "viterbi_tpc.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace dab;

static unsigned fnv(const void * p, size_t n)
{
  unsigned h = 2166136261u;
  const unsigned char * b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 16777619u;
  return h;
}

// The kernel of viterbi_kernels.cu with the 32 lanes of the warp emulated one after the other.
static void viterbi_emulated(const VitProfile & pr, const int16_t * soft, uint8_t * out)
{
  const int steps = pr.n_bits + 6;
  std::vector<unsigned> syms(steps, 0);
  for (int p = 0; p < 4 * steps; p++)
  {
    const int idx = vit_src_index(pr, p);
    int v = idx >= 0 ? soft[idx] : 0;
    v = (short)(v + 127); v = v < 0 ? 0 : (v > 255 ? 255 : v);
    syms[p / 4] |= (unsigned)v << (8 * (p & 3));
  }
  std::vector<unsigned long long> surv(steps);
  unsigned a[32], b[32], pk[32];
  for (int l = 0; l < 32; l++) { a[l] = (l == 0 ? 0u : 1000u) << 16; b[l] = 1000u << 16; }
  for (int t = 0; t < steps; t++)
  {
    unsigned w0 = 0, w1 = 0;
    for (int l = 0; l < 32; l++)
    {
      const unsigned x = syms[t] ^ vit_branch_mask(l);
      const unsigned metric = (x & 255) + ((x >> 8) & 255) + ((x >> 16) & 255) + (x >> 24);
      unsigned n0, n1; bool d0, d1;
      vit_acs(a[l], b[l], metric, n0, n1, d0, d1);
      w0 |= (unsigned)d0 << l; w1 |= (unsigned)d1 << l;
      pk[l] = (n0 >> 16) | (n1 & 0xffff0000u);
    }
    surv[t] = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
    for (int l = 0; l < 32; l++)
    {
      const unsigned va = pk[l >> 1], vb = pk[16 + (l >> 1)];
      a[l] = (l & 1) ? (va & 0xffff0000u) : (va << 16);
      b[l] = (l & 1) ? (vb & 0xffff0000u) : (vb << 16);
    }
  }
  unsigned r = 0;
  for (int e = steps - 1; e >= 6; e--) out[e - 6] = (uint8_t)vit_traceback_step(surv[e], r);
}

// The thread-per-code-word kernel (viterbi_tpc.cuh): the same host/device functions, one "thread".
static void viterbi_tpc_emulated(const VitProfile & pr, const int16_t * soft, uint8_t * out)
{
  const int steps = pr.n_bits + 6, rows = (steps + 4) / 5 * 5;
  std::vector<unsigned> syms(rows, 0);
  for (int t = 0; t < steps; t++)
  {
    int v[4];
    for (int g = 0; g < 4; g++) { const int idx = vit_src_index(pr, 4 * t + g); v[g] = idx >= 0 ? soft[idx] : 0; }
    syms[t] = tpc_pack_syms(v);
  }
  std::vector<unsigned long long> surv(rows);
  unsigned S[32];
  tpc_init(S);
  for (int t0 = 0; t0 < steps; t0 += 5)
  {
    const unsigned sy[5] = { syms[t0], syms[t0 + 1], syms[t0 + 2], syms[t0 + 3], syms[t0 + 4] };
    unsigned long long dec[5];
    tpc_five_steps(S, sy, dec, (t0 + 5) % TPC_RENORM == 0); // k_vit_tpc counts the blocks down instead
    for (int i = 0; i < 5; i++) surv[t0 + i] = dec[i];
    for (int i = 0; i < 32; i++) if ((S[i] & 0xffffu) > 60000u || (S[i] >> 16) > 60000u) { fprintf(stderr, "metric overflow\n"); exit(3); }
  }
  unsigned y = 0;
  for (int e = steps - 1; e >= 6; e--) out[e - 6] = (uint8_t)tpc_traceback_step(surv[e], (unsigned)tpc_decision_bit(e % 5, (int)y), y);
}

int main(int argc, char ** argv)
{
  if (argc >= 2 && !strcmp(argv[1], "addr"))
  {
    static const int uep_rates[] = { 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384 };
    std::vector<int32_t> addr(4 * 24 * 1024 + 24);
    for (int br : uep_rates)
      for (int lvl = 1; lvl <= 5; lvl++)
      {
        VitProfile p;
        if (!make_msc_profile(1, br, lvl, p)) { printf("ADDR 1 %d %d -1 0\n", lvl, br); continue; }
        const int n = profile_addresses(p, addr.data(), (int)addr.size());
        printf("ADDR 1 %d %d %d %u\n", lvl, br, n, fnv(addr.data(), sizeof(int32_t) * n));
      }
    for (int lvl = 0; lvl < 8; lvl++)
      for (int br = (lvl < 4 ? 8 : 32); br <= 192; br += (lvl < 4 ? 8 : 32))
      {
        VitProfile p;
        if (!make_msc_profile(0, br, lvl, p)) { printf("ADDR 0 %d %d -1 0\n", lvl, br); continue; }
        const int n = profile_addresses(p, addr.data(), (int)addr.size());
        printf("ADDR 0 %d %d %d %u\n", lvl, br, n, fnv(addr.data(), sizeof(int32_t) * n));
      }
    const VitProfile f = make_fic_profile();
    const int n = profile_addresses(f, addr.data(), (int)addr.size());
    printf("ADDR 9 0 0 %d %u\n", n, fnv(addr.data(), sizeof(int32_t) * n));
    return 0;
  }
  if (argc >= 4 && (!strcmp(argv[1], "vit") || !strcmp(argv[1], "vit2")))
  {
    const bool tpc = !strcmp(argv[1], "vit2");
    const int n_bits = atoi(argv[2]), count = atoi(argv[3]);
    const VitProfile pr = make_identity_profile(n_bits);
    std::vector<int16_t> soft(4 * (n_bits + 6));
    std::vector<uint8_t> bits(n_bits);
    for (int c = 0; c < count; c++)
    {
      if (fread(soft.data(), sizeof(int16_t), soft.size(), stdin) != soft.size()) return 2;
      if (tpc) viterbi_tpc_emulated(pr, soft.data(), bits.data());
      else viterbi_emulated(pr, soft.data(), bits.data());
      printf("VIT %d %d %u\n", n_bits, c, fnv(bits.data(), bits.size()));
    }
    return 0;
  }
  return 1;
}
