/* dab_tii.c — CPU restatement of the reference's TII detector (ofdm/tii_detector.cpp), part of the oracle: TEST
 * INFRASTRUCTURE, never linked into or called by the product path. Pinned against the reference's own TiiDetector object
 * (oracle/_ref, tests/test_oracle_vs_ref.py). The two constant tables are derived instead of copied:
 *   cMainIdPatternTable (tii_detector.cpp:19-90)  = the 70 bytes with four bits set, ascending (EN 300 401 table 42);
 *   cPhaseCorrTable (:92-125)                     = quadrant(PRS[k]) - quadrant(PRS[k+1]) mod 4 for the carrier pair k, k+1.
 * Every function cites the lines it follows. Float arithmetic as the reference compiles it (-ffast-math: |z| = sqrt(re^2 + im^2)). */
#include "dab_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { T_U = 2048, K_CARR = 1536, N_BLOCKS = 4, N_GROUPS = 8, GROUP = 24, BLOCK = 192, N_PAIRS = 768, N_MAIN = 70 };

typedef struct { float re, im; } c32;

typedef struct
{
  c32 null_sum[T_U];      /* mNullSymbolBufferVec */
  c32 decoded[N_PAIRS];   /* mDecodedBufferArr */
  int collisions, sub_id_coll;
  int carrier_delete;     /* mCarrierDelete = true */
  uint8_t pattern[N_MAIN];
  uint8_t phase_corr[N_PAIRS];
} tii_t;

static inline float cabs32(c32 v) { return sqrtf(v.re * v.re + v.im * v.im); }
static inline int fft_idx(int k) { return k < 0 ? k + T_U : k + 1; } /* fft_shift_skip_dc, glob_defs.h:284-287 */

void * dabo_tii_new(void)
{
  tii_t * t = (tii_t *)calloc(1, sizeof(*t));
  t->carrier_delete = 1;
  int n = 0;
  for (int b = 0; b < 256; b++) if (__builtin_popcount((unsigned)b) == 4) t->pattern[n++] = (uint8_t)b;
  float prs[2 * T_U];
  dabo_phase_table(prs);
  for (int i = 0; i < N_PAIRS; i++)
  {
    const int f = fft_idx(-K_CARR / 2 + 2 * i);
    /* PRS values are +-1 / +-j: quadrant 0..3 */
    const int qa = (int)lroundf(atan2f(prs[2 * f + 1], prs[2 * f]) / 1.57079632679f) & 3;
    const int qb = (int)lroundf(atan2f(prs[2 * (f + 1) + 1], prs[2 * (f + 1)]) / 1.57079632679f) & 3;
    t->phase_corr[i] = (uint8_t)((qa - qb) & 3);
  }
  return t;
}
void dabo_tii_free(void * h) { free(h); }
void dabo_tii_reset(void * h) /* :147-151 */
{
  tii_t * t = (tii_t *)h;
  memset(t->null_sum, 0, sizeof(t->null_sum));
  memset(t->decoded, 0, sizeof(t->decoded));
}
void dabo_tii_set_collisions(void * h, int on, int sub_id) { ((tii_t *)h)->collisions = on; ((tii_t *)h)->sub_id_coll = sub_id; }
void dabo_tii_tables(void * h, uint8_t pattern[70], uint8_t phase_corr[768])
{
  memcpy(pattern, ((tii_t *)h)->pattern, N_MAIN);
  memcpy(phase_corr, ((tii_t *)h)->phase_corr, N_PAIRS);
}

void dabo_tii_add(void * h, const float * fft2048) /* add_to_tii_buffer, :154-160 */
{
  tii_t * t = (tii_t *)h;
  for (int i = 0; i < T_U; i++) { t->null_sum[i].re += fft2048[2 * i]; t->null_sum[i].im += fft2048[2 * i + 1]; }
}

static c32 turn_phase(c32 v, int phase) /* :295-309 */
{
  switch (phase)
  {
  case 3: { c32 r = { -v.im, v.re }; return r; }
  case 2: { c32 r = { -v.re, -v.im }; return r; }
  case 1: { c32 r = { v.im, -v.re }; return r; }
  default: return v;
  }
}

static int cmp_strength(const void * a, const void * b)
{
  const float x = ((const dabo_tii_result *)a)->strength, y = ((const dabo_tii_result *)b)->strength;
  return x > y ? -1 : (x < y ? 1 : 0);
}

int dabo_tii_process(void * h, int threshold_db, dabo_tii_result * out, int cap) /* process_tii_data, :162-241 */
{
  tii_t * t = (tii_t *)h;
  /* _decode_and_accumulate_carrier_pairs, :247-267 */
  for (int k = -K_CARR / 2, i = 0; k < K_CARR / 2; k += 2, i++)
  {
    const int f = fft_idx(k);
    const c32 a = t->null_sum[f], b = t->null_sum[f + 1];
    const c32 prod = { a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im }; /* a conj(b) */
    t->decoded[i].re += 0.01f * (prod.re - t->decoded[i].re);
    t->decoded[i].im += 0.01f * (prod.im - t->decoded[i].im);
  }
  /* _collapse_tii_groups, :317-343 (with _remove_single_carrier_values, :269-293) */
  c32 buf[N_PAIRS];
  memcpy(buf, t->decoded, sizeof(buf));
  if (t->carrier_delete)
    for (int i = 0; i < BLOCK; i++)
    {
      float max = 0, sum = 0;
      int index = 0;
      for (int j = 0; j < N_BLOCKS; j++)
      {
        const float x = cabs32(buf[i + j * BLOCK]);
        sum += x;
        if (x > max) { max = x; index = j; }
      }
      const float min = (sum - max) / (N_BLOCKS - 1);
      if (sum < max * 1.5 && max > 0.0)
      {
        buf[i + index * BLOCK].re *= min / max;
        buf[i + index * BLOCK].im *= min / max;
      }
    }
  c32 etsi_c[BLOCK], non_c[BLOCK];
  float etsi_f[BLOCK], non_f[BLOCK];
  for (int i = 0; i < BLOCK; i++)
  {
    c32 e = { 0, 0 }, n = { 0, 0 };
    for (int b = 0; b < N_BLOCKS; b++)
    {
      const c32 x = buf[i + b * BLOCK];
      e.re += x.re; e.im += x.im;
      const c32 y = turn_phase(x, t->phase_corr[i + b * BLOCK]);
      n.re += y.re; n.im += y.im;
    }
    etsi_c[i] = e; non_c[i] = n;
  }
  /* _get_float_table_and_max_abs_value, :474-483 */
  float max = 0;
  for (int i = 0; i < BLOCK; i++) { etsi_f[i] = cabs32(etsi_c[i]); if (etsi_f[i] > max) max = etsi_f[i]; }
  for (int i = 0; i < BLOCK; i++) { non_f[i] = cabs32(non_c[i]); if (non_f[i] > max) max = non_f[i]; }
  /* _calculate_average_noise, :485-500 */
  float noise = 1e9f;
  for (int s = 0; s < GROUP; s++)
  {
    float avg = 0;
    for (int i = 0; i < N_GROUPS; i++) avg += etsi_f[s + i * GROUP];
    avg /= N_GROUPS;
    if (avg < noise) noise = avg;
  }

  int n_out = 0;
  dabo_tii_result res[GROUP * (1 + N_MAIN)];
  for (int sub = 0; sub < GROUP; sub++)
  {
    const float thr = noise * powf(10.0f, (float)threshold_db / 10.0f);
    /* _compare_etsi_and_non_etsi, :385-437 */
    c32 es = { 0, 0 }, ns = { 0, 0 };
    int ec = 0, nc = 0;
    unsigned ep = 0, np = 0;
    for (int i = 0; i < N_GROUPS; i++)
    {
      const int idx = sub + i * GROUP;
      if (etsi_f[idx] > thr) { ec++; ep |= 0x80u >> i; es.re += etsi_c[idx].re; es.im += etsi_c[idx].im; }
      if (non_f[idx] > thr) { nc++; np |= 0x80u >> i; ns.re += non_c[idx].re; ns.im += non_c[idx].im; }
    }
    int non_etsi = 0;
    if (ec >= 4 || nc >= 4) if (cabs32(ns) > cabs32(es)) non_etsi = 1;
    c32 sum = non_etsi ? ns : es;
    const int count = non_etsi ? nc : ec;
    const unsigned pattern = non_etsi ? np : ep;
    const c32 * ct = non_etsi ? non_c : etsi_c;
    const float * ft = non_etsi ? non_f : etsi_f;

    int main_id = 0;
    if (count == 4)
    {
      for (int m = 0; m < N_MAIN; m++) if (t->pattern[m] == pattern) { main_id = m; break; } /* _find_exact_main_id_match, :345-357 */
    }
    else if (count > 4)
    {
      /* _find_best_main_id_match, :359-383 */
      float max_level = 0;
      main_id = -1;
      sum.re = 0; sum.im = 0;
      for (int m = 0; m < N_MAIN; m++)
      {
        c32 val = { 0, 0 };
        for (int i = 0; i < N_GROUPS; i++)
          if (t->pattern[m] & (0x80u >> i)) { val.re += ct[sub + GROUP * i].re; val.im += ct[sub + GROUP * i].im; }
        if (cabs32(val) > max_level) { max_level = cabs32(val); sum = val; main_id = m; }
      }
    }
    if (count >= 4)
    {
      dabo_tii_result e = { main_id & 0xff, sub, cabs32(sum) / max / 4, atan2f(sum.im, sum.re) * 57.29577951308232f, non_etsi };
      res[n_out++] = e;
    }
    if (count > 4 && t->collisions)
    {
      /* _find_collisions, :439-472 */
      c32 s2 = { 0, 0 };
      for (int i = 0; i < N_GROUPS; i++)
        if (!(t->pattern[main_id] & (0x80u >> i)))
        {
          const int idx = sub + GROUP * i;
          if (ft[idx] > thr) { s2.re += ct[idx].re; s2.im += ct[idx].im; }
        }
      const float strength = cabs32(s2) / max / (float)(count - 4), ph = atan2f(s2.im, s2.re) * 57.29577951308232f;
      if (sub == t->sub_id_coll)
      {
        for (int m = 0; m < N_MAIN; m++)
        {
          const unsigned p2 = t->pattern[m] & pattern;
          if (__builtin_popcount(p2) == 4 && m != main_id)
          {
            dabo_tii_result e = { m, sub, strength, ph, non_etsi };
            res[n_out++] = e;
          }
        }
      }
      else
      {
        dabo_tii_result e = { 99, sub, strength, ph, non_etsi };
        res[n_out++] = e;
      }
    }
  }
  memset(t->null_sum, 0, sizeof(t->null_sum)); /* _reset_null_symbol_buffer */
  qsort(res, (size_t)n_out, sizeof(res[0]), cmp_strength); /* std::sort by strength, descending (order of equal strengths unspecified) */
  for (int i = 0; i < n_out && i < cap; i++) out[i] = res[i];
  return n_out;
}

void dabo_tii_decoded(void * h, float * out_re_im) { memcpy(out_re_im, ((tii_t *)h)->decoded, sizeof(c32) * N_PAIRS); }
