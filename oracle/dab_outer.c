/* dab_outer.c — CPU restatement of the DAB+ outer code (TEST INFRASTRUCTURE, part of libdab_oracle.so).
 *
 *   Galois / ReedSolomon(8, 0435, 0, 1, 10).dec(.., 135)   backend/galois.cpp:36-143, backend/reed_solomon.cpp:140-440
 *   FirecodeChecker::check / check_and_correct_6bits        backend/firecode_checker.cpp:41-184, firecode_checker.h:55-72
 *   Mp4Processor::add_to_frame / _process_reed_solomon_frame / _process_super_frame (header, AU borders, AU CRC)
 *                                                            backend/audio/mp4processor.cpp:95-241, 246-334
 * The Reed-Solomon and Fire-code parts are pinned against the reference's own objects (dabref_rs_decode, dabref_firecode_*);
 * the super-frame state machine lives in a Qt/faad class that cannot be built here and is restated from its source.
 */
#include "dab_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- GF(2^8), generator polynomial 0435 (octal) = 0x11D; index 255 stands for log(0) */
static uint8_t gf_exp[256], gf_log[256];
static int gf_ready = 0;
static void gf_init(void)
{
  if (gf_ready) return;
  unsigned sr = 1;
  gf_log[0] = 255; gf_exp[255] = 0;
  for (int i = 0; i < 255; i++)
  {
    gf_log[sr] = (uint8_t)i; gf_exp[i] = (uint8_t)sr;
    sr <<= 1;
    if (sr & 0x100) sr ^= 0x11D;
    sr &= 255;
  }
  gf_ready = 1;
}
static inline int modnn(int x) { while (x >= 255) { x -= 255; x = (x >> 8) + (x & 255); } return x; }
static inline unsigned gf_mul_poly(unsigned a, unsigned b) { return (a == 0 || b == 0) ? 0 : gf_exp[modnn(gf_log[a] + gf_log[b])]; }
static inline unsigned gf_div_poly(unsigned a, unsigned b) { return a == 0 ? 0 : gf_exp[modnn(255 + gf_log[a] - gf_log[b])]; }
static inline int pow_power(int a, int n) { return a == 0 ? 0 : (a * n) % 255; }

/* ReedSolomon::decode_rs on the 255-symbol block (reed_solomon.cpp:160-260); returns what dec() returns */
static int rs_decode_block(uint8_t * data)
{
  enum { NROOTS = 10, NN = 255 };
  uint8_t syn[NROOTS], lambda[NROOTS + 1], root_tab[NROOTS], loc_tab[NROOTS], omega[NROOTS + 1];
  /* syndromes by Horner, root i = alpha^i (fcr 0, prim 1) */
  unsigned any = 0;
  for (int i = 0; i < NROOTS; i++)
  {
    unsigned s = data[0];
    for (int j = 1; j < NN; j++)
      s = s == 0 ? data[j] : (data[j] ^ gf_exp[modnn(gf_log[s] + pow_power(modnn(0 + i), 1))]);
    syn[i] = (uint8_t)s;
    any |= s;
  }
  if (any == 0) return 0;
  /* Berlekamp-Massey (computeLambda), lambda ends in power form */
  int deg_lambda = 0;
  {
    uint8_t corr[NROOTS], old[NROOTS];
    unsigned K = 1, L = 0;
    memset(corr, 0, sizeof(corr));
    memset(lambda, 0, sizeof(lambda));
    unsigned error = syn[0];
    lambda[0] = 1;
    corr[1] = 1;
    while (K < NROOTS)
    {
      memcpy(old, lambda, NROOTS);
      for (int i = 0; i < NROOTS; i++) lambda[i] ^= (uint8_t)gf_mul_poly(error, corr[i]);
      if (2 * L < K && error != 0)
      {
        L = K - L;
        for (int i = 0; i < NROOTS; i++) corr[i] = (uint8_t)gf_div_poly(old[i], error);
      }
      for (int i = NROOTS - 1; i >= 1; i--) corr[i] = corr[i - 1];
      corr[0] = 0;
      error = syn[K];
      for (unsigned i = 1; i <= K; i++) error ^= gf_mul_poly(syn[K - i], lambda[i]);
      K++;
    }
    for (int i = 0; i < NROOTS; i++) lambda[i] ^= (uint8_t)gf_mul_poly(error, corr[i]);
    for (int i = 0; i < NROOTS; i++) { if (lambda[i] != 0) deg_lambda = i; lambda[i] = gf_log[lambda[i]]; }
  }
  /* Chien search (computeErrors): iprim = 1 */
  int root_count = 0;
  {
    uint8_t work[NROOTS + 1];
    memcpy(work, lambda, NROOTS + 1);
    for (int i = 1, k = 0; i <= NN; i++, k++)
    {
      unsigned result = 1;
      for (int j = deg_lambda; j > 0; j--)
        if (work[j] != NN)
        {
          work[j] = (uint8_t)modnn(work[j] + j);
          result ^= gf_exp[work[j]];
        }
      if (result != 0) continue;
      root_tab[root_count] = (uint8_t)i;
      loc_tab[root_count] = (uint8_t)k;
      root_count++;
    }
    if (root_count != deg_lambda) return -1;
  }
  /* error evaluator (computeOmega) */
  int deg_omega = 0;
  for (int i = 0; i < NROOTS; i++)
  {
    unsigned tmp = 0;
    for (int j = deg_lambda < i ? deg_lambda : i; j >= 0; j--)
      if (gf_log[syn[i - j]] != NN && lambda[j] != NN) tmp ^= gf_exp[modnn(gf_log[syn[i - j]] + lambda[j])];
    if (tmp != 0) deg_omega = i;
    omega[i] = gf_log[tmp];
  }
  omega[NROOTS] = NN;
  /* Forney (decode_rs tail) */
  for (int j = root_count - 1; j >= 0; j--)
  {
    unsigned num1 = 0;
    for (int i = deg_omega; i >= 0; i--)
      if (omega[i] != NN) num1 ^= gf_exp[modnn(omega[i] + pow_power(i, root_tab[j]))];
    const unsigned num2 = gf_exp[modnn(pow_power(root_tab[j], modnn(255 + 0 - 1)) + NN)];
    unsigned den = 0;
    for (int i = (deg_lambda < NROOTS - 1 ? deg_lambda : NROOTS - 1) & ~1; i >= 0; i -= 2)
      if (lambda[i + 1] != NN) den ^= gf_exp[modnn(lambda[i + 1] + pow_power(i, root_tab[j]))];
    if (den == 0) return -1;
    if (num1 != 0)
    {
      if (loc_tab[j] >= NN - NROOTS) root_count--;
      else
      {
        const int t1 = NN - gf_log[den];
        int t2 = modnn(gf_log[num1] + gf_log[num2]);
        t2 = modnn(t2 + t1);
        data[loc_tab[j]] ^= gf_exp[t2];
      }
    }
  }
  return root_count;
}

int dabo_rs_decode(const uint8_t * in120, uint8_t * out110)
{
  gf_init();
  uint8_t rf[255];
  memset(rf, 0, 135);
  memcpy(rf + 135, in120, 120);
  const int ret = rs_decode_block(rf);
  memcpy(out110, rf + 135, 110);
  return ret;
}

/* systematic encoder for test vectors: parity = data(x) x^10 mod g(x), g = prod (x - alpha^i), i = 0..9 */
void dabo_rs_encode(const uint8_t * in110, uint8_t * out120)
{
  gf_init();
  uint8_t g[11];
  memset(g, 0, sizeof(g));
  g[0] = 1;
  for (int i = 0; i < 10; i++)
  {
    for (int j = i + 1; j > 0; j--) g[j] = (uint8_t)(g[j - 1] ^ gf_mul_poly(g[j], gf_exp[i]));
    g[0] = (uint8_t)gf_mul_poly(g[0], gf_exp[i]);
  }
  uint8_t par[10];
  memset(par, 0, sizeof(par));
  for (int i = 0; i < 110; i++)
  {
    const unsigned fb = in110[i] ^ par[0];
    memmove(par, par + 1, 9);
    par[9] = 0;
    if (fb) for (int j = 0; j < 10; j++) par[j] ^= (uint8_t)gf_mul_poly(fb, g[9 - j]);
  }
  memcpy(out120, in110, 110);
  memcpy(out120 + 110, par, 10);
}

/* ---- Fire code (firecode_checker.cpp): g(x) = (x^11+1)(x^5+x^3+x^2+x+1) = 0x782f, 16 parity bits in x[0..1] over x[2..10] */
static uint16_t fc_tab[256], fc_syn[65536];
static int fc_ready = 0;
static uint16_t fc_crc(const uint8_t * x)
{
  uint16_t crc = 0;
  for (int i = 2; i < 11; i++) crc = (uint16_t)((crc << 8) ^ fc_tab[(crc >> 8) ^ x[i]]);
  for (int i = 0; i < 2; i++) crc = (uint16_t)((crc << 8) ^ fc_tab[(crc >> 8) ^ x[i]]);
  return crc;
}
static const uint8_t FC_PATTERN[124] = {
   17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 30, 31, 34, 36, 38, 40, 42, 44, 46, 50, 52, 54, 56, 60, 62, 68, 72, 76,
   84, 88, 92,100,104,108,120,124,136,152,168,184,200,216,248,
   33, 35, 37, 39, 41, 43, 45, 49, 51, 53, 55, 57, 59, 61, 63,
   66, 70, 74, 78, 82, 86, 90, 98,102,106,110,114,118,122,126,
  132,140,148,156,164,172,180,196,204,212,220,228,236,244,252,
    1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15, 16, 29, 32, 48, 58, 64, 80, 96,112,116,128,144,160,176,192,
  208,224,232,240 };
static void fc_init(void)
{
  if (fc_ready) return;
  for (int i = 0; i < 256; i++)
  {
    uint16_t crc = (uint16_t)(i << 8);
    for (int j = 0; j < 8; j++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x782f) : (uint16_t)(crc << 1);
    fc_tab[i] = crc;
  }
  memset(fc_syn, 0, sizeof(fc_syn));
  uint8_t e[11];
  memset(e, 0, 11);
  /* the table keeps the FIRST burst that produces a syndrome, in exactly this order (firecode_checker.cpp:66-140) */
  static const struct { int shift, first, last, bytes; } pass[4] = { { 0, 0, 124, 11 }, { 4, 0, 45, 10 }, { 2, 45, 75, 10 }, { 6, 60, 90, 10 } };
  for (int p = 0; p < 4; p++)
    for (int i = 0; i < pass[p].bytes; i++)
      for (int j = pass[p].first; j < pass[p].last; j++)
      {
        const int sh = pass[p].shift, bit = i * 8 + sh;
        if (sh == 0) e[i] = FC_PATTERN[j];
        else { e[i] = (uint8_t)(FC_PATTERN[j] >> sh); e[i + 1] = (uint8_t)(FC_PATTERN[j] << (8 - sh)); }
        const uint16_t s = fc_crc(e);
        if (fc_syn[s] == 0) fc_syn[s] = (uint16_t)((bit << 8) + FC_PATTERN[j]);
        e[i] = 0;
        if (sh != 0) e[i + 1] = 0;
      }
  fc_ready = 1;
}
int dabo_firecode_check(const uint8_t * x11) { fc_init(); return fc_crc(x11) == 0; }
int dabo_firecode_check_and_correct(uint8_t * x)
{
  fc_init();
  const uint16_t s = fc_crc(x);
  if (s == 0) return 1;
  const uint8_t err = (uint8_t)(fc_syn[s] & 0xff);
  if (!err) return 0;
  const int bit = fc_syn[s] >> 8;
  x[bit / 8] ^= (uint8_t)(err >> (bit % 8));
  x[bit / 8 + 1] ^= (uint8_t)(err << (8 - (bit % 8)));
  return 1;
}
/* the 65536-entry syndrome table itself (the product builds its own copy; compared in the tests) */
void dabo_firecode_syndrome_table(uint16_t * out) { fc_init(); memcpy(out, fc_syn, sizeof(fc_syn)); }

/* check_crc_bytes (backend/crc.cpp:89-96) */
int dabo_check_crc_bytes(const uint8_t * msg, int len)
{
  const uint16_t acc = dabo_calc_crc(msg, len);
  return (uint16_t)((msg[len] << 8) | msg[len + 1]) == acc;
}

/* ---- Mp4Processor: super-frame synchronisation and dissection over a run of logical frames (mp4processor.cpp:95-334).
 * frame_bits: n_frames x 24*bit_rate bytes, one bit each (what Backend hands to FrameProcessor::add_to_frame).
 * out: one record per super-frame the processor ATTEMPTED (in order); payload: 110*(bit_rate/8) bytes per record (mOutVec). */
int dabo_dabplus_run(const uint8_t * frame_bits, int bit_rate, int n_frames, dabo_superframe * out, int cap, uint8_t * payload)
{
  gf_init();
  fc_init();
  const int num_bytes = 3 * bit_rate, rs_dims = bit_rate / 8, sf_bytes = 110 * rs_dims;
  uint8_t * ring = (uint8_t *)calloc((size_t)rs_dims * 120, 1);
  uint8_t * outv = (uint8_t *)calloc((size_t)sf_bytes, 1);
  int fill = 0, in_buf = 0, sync = 0, n_out = 0;
  for (int f = 0; f < n_frames; f++)
  {
    for (int i = 0; i < num_bytes; i++)
    {
      unsigned t = 0;
      for (int j = 0; j < 8; j++) t = (t << 1) | (frame_bits[(size_t)f * 24 * bit_rate + 8 * i + j] & 1u);
      ring[fill * num_bytes + i] = (uint8_t)t;
    }
    in_buf++;
    fill = (fill + 1) % 5;
    if (in_buf < 5) continue;
    if (sync == 0)
    {
      if (dabo_firecode_check(&ring[fill * num_bytes])) sync = 4;
      else in_buf = 4;
    }
    if (!sync) continue;
    in_buf = 0;
    /* _process_super_frame */
    dabo_superframe rec;
    memset(&rec, 0, sizeof(rec));
    rec.first_frame = f - 4;
    const int base = fill * num_bytes;
    for (int j = 0; j < rs_dims; j++)
    {
      uint8_t in[120], o[110];
      for (int k = 0; k < 120; k++) in[k] = ring[(base + j + k * rs_dims) % (rs_dims * 120)];
      const int ler = dabo_rs_decode(in, o);
      if (ler < 0) rec.rs_errors++; else rec.rs_corrections += ler;
      for (int k = 0; k < 110; k++) outv[j + k * rs_dims] = o[k];
    }
    int ok = dabo_firecode_check_and_correct(outv);
    if (ok) { uint8_t first[11]; for (int k = 0; k < 11; k++) first[k] = ring[(base + k) % (rs_dims * 120)]; rec.fc_corrected = memcmp(outv, first, 11) != 0; }
    rec.ok = ok;
    if (ok)
    {
      rec.dac_rate = (outv[2] >> 6) & 1; rec.sbr_flag = (outv[2] >> 5) & 1; rec.aac_channel_mode = (outv[2] >> 4) & 1;
      rec.ps_flag = (outv[2] >> 3) & 1; rec.mpeg_surround = outv[2] & 7;
      int * a = rec.au_start;
      switch (2 * rec.dac_rate + rec.sbr_flag)
      {
      case 0: rec.num_aus = 4; a[0] = 8; a[1] = outv[3] * 16 + (outv[4] >> 4); a[2] = (outv[4] & 0xf) * 256 + outv[5]; a[3] = outv[6] * 16 + (outv[7] >> 4); a[4] = sf_bytes; break;
      case 1: rec.num_aus = 2; a[0] = 5; a[1] = outv[3] * 16 + (outv[4] >> 4); a[2] = sf_bytes; break;
      case 2: rec.num_aus = 6; a[0] = 11; a[1] = outv[3] * 16 + (outv[4] >> 4); a[2] = (outv[4] & 0xf) * 256 + outv[5]; a[3] = outv[6] * 16 + (outv[7] >> 4);
              a[4] = (outv[7] & 0xf) * 256 + outv[8]; a[5] = outv[9] * 16 + (outv[10] >> 4); a[6] = sf_bytes; break;
      default: rec.num_aus = 3; a[0] = 6; a[1] = outv[3] * 16 + (outv[4] >> 4); a[2] = (outv[4] & 0xf) * 256 + outv[5]; a[3] = sf_bytes; break;
      }
      for (int u = 0; u < rec.num_aus; u++)
      {
        const int len = a[u + 1] - a[u] - 2;
        /* 0: length check failed (mp4processor.cpp:326), 1: CRC good, 2: CRC error. An AU that would leave the super-frame cannot be
         * CRC-checked without reading beyond mOutVec (the reference does read there); it is reported as a CRC error */
        if (len > 960 || len < 0) rec.au_state[u] = 0;
        else if (a[u] + len + 2 > sf_bytes) rec.au_state[u] = 2;
        else rec.au_state[u] = dabo_check_crc_bytes(&outv[a[u]], len) ? 1 : 2;
      }
      sync = 4;
    }
    else
    {
      sync--;
      if (sync == 0) in_buf = 4;
    }
    if (n_out < cap)
    {
      out[n_out] = rec;
      if (payload) memcpy(payload + (size_t)n_out * sf_bytes, outv, (size_t)sf_bytes);
    }
    n_out++;
  }
  free(ring); free(outv);
  return n_out;
}
