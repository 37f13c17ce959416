/* dab_oracle.c — CPU restatement of the reference's DAB Mode-I hot path. TEST INFRASTRUCTURE ONLY.
 *
 * See dab_oracle.h. Citations are file:line under /root/reference/src. Integer stages (tables,
 * depuncturing, time de-interleaving, Viterbi, PRBS, CRC, control counters) are bit-exact restatements;
 * floating-point stages follow the reference's order of operations in float32 (the reference itself is
 * built with -ffast-math, so its last bits are compiler dependent; tests allow the tolerance stated in
 * BASELINE.json for them).
 */
#include "dab_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * Mode-I constants (common/glob_defs.h:40-55, common/dab_constants.h:55-65)
 * ---------------------------------------------------------------------------------------------- */
enum
{
  L_SYMS = 76, K_CARR = 1536, T_NULL = 2656, T_FRAME = 196608, T_S = 2552, T_U = 2048, T_G = 504,
  CARR_DIFF = 1000, BITS_PER_SYM = 3072, FIC_VIT_IN = 2304, FIC_VIT_OUT = 768, FIB_BITS = 256,
  FS = 2048000, CIF_BITS = 55296
};
#define PI_F ((float)M_PI)
#define PI_2_F ((float)(M_PI / 2.0))
#define PI_4_F ((float)(M_PI / 4.0))
#define TWO_PI_F ((float)(2.0 * M_PI))
#define RAD_PER_DEG_F ((float)(M_PI / 180.0))

typedef struct { float re, im; } cf32;

static inline cf32 cmul(cf32 a, cf32 b) { cf32 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline cf32 cconj(cf32 a) { cf32 r = { a.re, -a.im }; return r; }
static inline float cabs32(cf32 a) { return sqrtf(a.re * a.re + a.im * a.im); }
static inline float clampf_sym(float v, float lim) { return v > lim ? lim : (v < -lim ? -lim : v); }

/* ------------------------------------------------------------------------------------------------
 * Frequency interleaver (ofdm/freq_interleaver.cpp:40-76): Pi(0)=0, Pi(i)=(13 Pi(i-1)+511) mod 2048,
 * keep 256..1792 except 1024, subtract 1024.
 * ---------------------------------------------------------------------------------------------- */
void dabo_freq_interleaver(int16_t out[1536])
{
  int n = 0;
  int v = 0;
  for (int i = 0; i < T_U; i++)
  {
    if (i > 0) v = (13 * v + 511) % T_U;
    if (v == T_U / 2 || v < 256 || v > 256 + K_CARR) continue;
    out[n++] = (int16_t)(v - T_U / 2);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Phase reference symbol (ofdm/phasetable.cpp:35-135, ETSI EN 300 401 table 44 + h matrix):
 * 48 blocks of 32 carriers, phi_k = pi/2 * (h[row][k - kmin] + n).
 * ---------------------------------------------------------------------------------------------- */
static const uint8_t PRS_ROW[48] = { 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3, 0,1,2,3,
                                     0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1, 0,3,2,1 };
static const uint8_t PRS_N[48] = { 1,2,0,1, 3,2,2,3, 2,1,2,3, 1,2,3,3, 2,2,2,1, 1,3,1,2,
                                   3,1,1,1, 2,2,1,0, 2,2,3,3, 0,2,1,3, 3,3,3,0, 3,0,1,1 };
static const uint8_t PRS_H[4][32] = {
  { 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1, 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1 },
  { 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0, 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0 },
  { 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3, 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3 },
  { 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2, 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2 } };

static int prs_quadrant(int k) /* k in [-768,-1] or [1,768] */
{
  const int blk = k < 0 ? (k + 768) / 32 : 24 + (k - 1) / 32;
  const int kmin = k < 0 ? -768 + 32 * blk : 1 + 32 * (blk - 24);
  return PRS_H[PRS_ROW[blk]][k - kmin] + PRS_N[blk];
}

static void build_phase_table(cf32 * t)
{
  memset(t, 0, sizeof(cf32) * T_U);
  for (int k = 1; k <= K_CARR / 2; k++)
  {
    const float pp = PI_2_F * (float)prs_quadrant(k);
    const float pm = PI_2_F * (float)prs_quadrant(-k);
    t[k].re = cosf(pp); t[k].im = sinf(pp);
    t[T_U - k].re = cosf(pm); t[T_U - k].im = sinf(pm);
  }
}

void dabo_phase_table(float out[4096]) { build_phase_table((cf32 *)out); }

/* ------------------------------------------------------------------------------------------------
 * FFT: the reference calls FFTW3f (third party, not vendored, unpinned in vcpkg.json; call sites
 * main/dab_processor.cpp:63,201,276,338 and ofdm/phasereference.cpp:51-52). Restated as its
 * published definition: unnormalised DFT, X[k] = sum x[n] e^{sign j 2 pi k n / N}; evaluated by a
 * radix-2 decimation in time in double, rounded to float once at the end.
 * ---------------------------------------------------------------------------------------------- */
static double g_tw_re[T_U / 2], g_tw_im[T_U / 2];
static int g_rev[T_U];
static int g_fft_ready = 0;

static void fft_init(void)
{
  if (g_fft_ready) return;
  for (int k = 0; k < T_U / 2; k++)
  {
    const double a = 2.0 * M_PI * (double)k / (double)T_U;
    g_tw_re[k] = cos(a);
    g_tw_im[k] = sin(a);
  }
  for (int i = 0; i < T_U; i++)
  {
    int r = 0;
    for (int b = 0; b < 11; b++) if (i & (1 << b)) r |= 1 << (10 - b);
    g_rev[i] = r;
  }
  g_fft_ready = 1;
}

static void fft2048(const cf32 * in, cf32 * out, int sign)
{
  static double wr[T_U], wi[T_U];
  fft_init();
  for (int i = 0; i < T_U; i++) { wr[g_rev[i]] = in[i].re; wi[g_rev[i]] = in[i].im; }
  for (int len = 2; len <= T_U; len <<= 1)
  {
    const int half = len / 2, step = T_U / len;
    for (int s = 0; s < T_U; s += len)
    {
      for (int k = 0; k < half; k++)
      {
        const double tr = g_tw_re[k * step], ti = (double)sign * g_tw_im[k * step];
        const double xr = wr[s + k + half] * tr - wi[s + k + half] * ti;
        const double xi = wr[s + k + half] * ti + wi[s + k + half] * tr;
        wr[s + k + half] = wr[s + k] - xr; wi[s + k + half] = wi[s + k] - xi;
        wr[s + k] += xr; wi[s + k] += xi;
      }
    }
  }
  for (int i = 0; i < T_U; i++) { out[i].re = (float)wr[i]; out[i].im = (float)wi[i]; }
}

void dabo_fft2048(const float * in, float * out, int sign) { fft2048((const cf32 *)in, (cf32 *)out, sign); }

/* ------------------------------------------------------------------------------------------------
 * Sample formats. u8: (v - 127.38)/128 (devices/filereaders/raw_files/raw_reader.cpp:66-70);
 * i16: v / 32768 (xml_reader.cpp:254-372 with 16 bit container, wav_reader.cpp:164).
 * ---------------------------------------------------------------------------------------------- */
void dabo_convert_u8(const uint8_t * in, float * out, int64_t n_samples)
{
  for (int64_t i = 0; i < 2 * n_samples; i++) out[i] = ((float)in[i] - 127.38f) / 128.0f;
}
void dabo_convert_i16(const int16_t * in, float * out, int64_t n_samples)
{
  for (int64_t i = 0; i < 2 * n_samples; i++) out[i] = (float)in[i] / 32768.0f;
}

/* xml_reader.cpp:43-51: scaler = 2^(bits-1) computed in i32, so 32 bits per channel give -2^31 (the x86 build wraps):
 * int32 files come out negated, which the differential demodulation does not notice */
static float xml_scaler(int bits)
{
  uint32_t r = 1;
  while (--bits > 0) r <<= 1;
  return (float)(int32_t)r;
}

/* one value of an XML/UFF file, xml_reader.cpp:254-372 (the same expressions repeat in the QI / I / Q readers) */
static float xml_value(const uint8_t * p, int container, int msb, float scaler, int order)
{
  switch (container)
  {
  case 0: return order == 0 ? (float)((int8_t)p[0]) / 127.0f : (float)((int8_t)p[0] / 127.0); /* :266 float, :411,:560 double */
  case 1: return ((float)p[0] - 127.38f) / 128.0f;                                             /* mapTable, :93-96 */
  case 2:
  {
    const int16_t v = msb ? (int16_t)((p[0] << 8) | p[1]) : (int16_t)((p[1] << 8) | p[0]);     /* :286-301 */
    return (float)v / scaler;
  }
  case 3:
  {
    int32_t v = msb ? ((p[0] << 16) | (p[1] << 8) | p[2]) : ((p[2] << 16) | (p[1] << 8) | p[0]); /* :308-341 */
    if (v & 0x800000) v |= (int32_t)0xFF000000;
    return (float)v / scaler;
  }
  case 4:
  {
    const uint32_t u = msb ? (((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]) : (((uint32_t)p[3] << 24) | (p[2] << 16) | (p[1] << 8) | p[0]);
    return (float)(int32_t)u / scaler;                                                          /* :346-366 */
  }
  default:
  {
    union { uint32_t u; float f; } c;                                                           /* UCnv, :376-397 */
    c.u = msb ? (((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]) : (((uint32_t)p[3] << 24) | (p[2] << 16) | (p[1] << 8) | p[0]);
    return c.f;
  }
  }
}

int dabo_convert_samples(const uint8_t * in, int container, int bits, int msb_first, int order, int64_t n_samples, float * out)
{
  static const int width[6] = { 1, 1, 2, 3, 4, 4 };
  if (container == 6 || container == 7)
  {
    /* RIFF/WAVE through libsndfile's sf_readf_float (wav_reader.cpp:164; libsndfile is not part of the reference tree):
     * 8-bit PCM (v - 128) / 128, 32-bit PCM v / 2^31, little endian, I then Q */
    if (order != 0) return -1;
    for (int64_t i = 0; i < 2 * n_samples; i++)
    {
      if (container == 6) out[i] = (float)((int)in[i] - 128) / 128.0f;
      else
      {
        const uint8_t * p = in + 4 * i;
        const uint32_t u = ((uint32_t)p[3] << 24) | (p[2] << 16) | (p[1] << 8) | p[0];
        out[i] = (float)(int32_t)u / 2147483648.0f;
      }
    }
    return 0;
  }
  if (container < 0 || container > 5 || order < 0 || order > 3) return -1;
  const int b = width[container];
  const float scaler = xml_scaler(bits > 0 ? bits : 8 * b);
  for (int64_t i = 0; i < n_samples; i++)
  {
    if (order <= 1)
    {
      const float x = xml_value(in + (2 * i) * b, container, msb_first, scaler, order);
      const float y = xml_value(in + (2 * i + 1) * b, container, msb_first, scaler, order);
      out[2 * i] = order == 0 ? x : y;
      out[2 * i + 1] = order == 0 ? y : x;
    }
    else
    {
      const float x = xml_value(in + i * b, container, msb_first, scaler, order);
      out[2 * i] = order == 2 ? x : 0.0f;
      out[2 * i + 1] = order == 2 ? 0.0f : x;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SampleReader's optional DC removal and IQ-imbalance correction, scalar build (ofdm/sample_reader.cpp:216-243;
 * members sample_reader.h:102-106; mean_filter glob_defs.h:217-220), sample by sample in float as the reference runs it.
 * state: { meanI, meanQ, meanII, meanQQ, meanIQ }, in and out.
 * ---------------------------------------------------------------------------------------------- */
void dabo_dc_iq_correct(const float * in_re_im, int64_t n, int do_iq, float state[5], float * out_re_im)
{
  float mean_i = state[0], mean_q = state[1], mean_ii = state[2], mean_qq = state[3], mean_iq = state[4];
  const float alpha = 1.0f / (float)FS / 1.00f;
  for (int64_t i = 0; i < n; i++)
  {
    const float v_i = in_re_im[2 * i], v_q = in_re_im[2 * i + 1];
    mean_i += alpha * (v_i - mean_i);
    mean_q += alpha * (v_q - mean_q);
    if (do_iq)
    {
      const float x_i = v_i - mean_i;
      const float x_q = v_q - mean_q;
      mean_ii += alpha * (x_i * x_i - mean_ii);
      mean_iq += alpha * (x_i * x_q - mean_iq);
      const float phi = mean_iq / mean_ii;
      const float x_q_corr = x_q - phi * x_i;
      mean_qq += alpha * (x_q_corr * x_q_corr - mean_qq);
      const float gain_q = sqrtf(mean_ii / mean_qq);
      out_re_im[2 * i] = x_i;
      out_re_im[2 * i + 1] = x_q_corr * gain_q;
    }
    else
    {
      out_re_im[2 * i] = v_i - mean_i;
      out_re_im[2 * i + 1] = v_q - mean_q;
    }
  }
  state[0] = mean_i; state[1] = mean_q; state[2] = mean_ii; state[3] = mean_qq; state[4] = mean_iq;
}

/* ------------------------------------------------------------------------------------------------
 * Sample-rate conversion of the file readers when the reference is built without liquid-dsp (the default): blocks of
 * rate / 1000 input samples become 2048 output samples by linear interpolation. Restated loop for loop; the readers
 * themselves need Qt (and libsndfile), so this boundary is pinned by the source text only.
 * ---------------------------------------------------------------------------------------------- */
/* XmlReader ctor tables (xml_reader.cpp:70-76) + readSamples (:212-231). Returns the number of output samples. */
int64_t dabo_resample_xml(const float * in_re_im, int64_t n_in, int sample_rate, float * out_re_im)
{
  const cf32 * in = (const cf32 *)in_re_im;
  cf32 * out = (cf32 *)out_re_im;
  int16_t map_int[2048];
  float map_float[2048];
  for (int i = 0; i < 2048; i++)
  {
    const float in_val = (float)(sample_rate / 1000);
    map_int[i] = (int16_t)(floor(i * (in_val / 2048.0)));
    map_float[i] = i * (in_val / 2048.0f) - map_int[i];
  }
  const int conv_size = sample_rate / 1000;
  cf32 * conv = (cf32 *)calloc((size_t)conv_size + 1, sizeof(cf32)); /* convBuffer.resize(convBufferSize + 1): zeros */
  int64_t n_out = 0;
  for (int64_t pos = 0; pos + conv_size <= n_in; pos += conv_size)
  {
    memcpy(&conv[1], &in[pos], sizeof(cf32) * (size_t)conv_size); /* the reader fills &convBuffer[1] */
    for (int i = 0; i < 2048; i++)
    {
      const int16_t base = map_int[i];
      const float ratio = map_float[i];
      out[n_out].re = conv[base + 1].re * ratio + conv[base].re * (1.0f - ratio);
      out[n_out].im = conv[base + 1].im * ratio + conv[base].im * (1.0f - ratio);
      n_out++;
    }
    conv[0] = conv[conv_size];
  }
  free(conv);
  return n_out;
}

/* WavReader ctor tables (wav_reader.cpp:66-79) + run (:196-211) */
int64_t dabo_resample_wav(const float * in_re_im, int64_t n_in, int sample_rate, float * out_re_im)
{
  const cf32 * in = (const cf32 *)in_re_im;
  cf32 * out = (cf32 *)out_re_im;
  int16_t map_int[2048];
  float map_float[2048];
  const int16_t conv_size = (int16_t)(sample_rate / 1000);
  for (int i = 0; i < 2048; i++)
  {
    const float in_val = (float)sample_rate / 1000.0f;
    map_int[i] = (int16_t)(floorf((float)i * (in_val / 2048.0f)));
    map_float[i] = (float)i * (in_val / 2048.0f) - (float)map_int[i];
  }
  cf32 * conv = (cf32 *)calloc((size_t)conv_size + 1, sizeof(cf32));
  int conv_index = 0;
  int64_t n_out = 0;
  for (int64_t i = 0; i < n_in; i++)
  {
    conv[conv_index++] = in[i];
    if (conv_index > conv_size)
    {
      for (int j = 0; j < 2048; j++)
      {
        const int16_t base = map_int[j];
        const float ratio = map_float[j];
        out[n_out].re = conv[base + 1].re * ratio + conv[base].re * (1 - ratio);
        out[n_out].im = conv[base + 1].im * ratio + conv[base].im * (1 - ratio);
        n_out++;
      }
      conv[0] = conv[conv_size];
      conv_index = 1;
    }
  }
  free(conv);
  return n_out;
}

/* ------------------------------------------------------------------------------------------------
 * Energy dispersal PRBS x^9 + x^5 + 1, all-ones start (decoder/fic_decoder.cpp:59-73,
 * backend/backend.cpp:72-84).
 * ---------------------------------------------------------------------------------------------- */
void dabo_prbs(uint8_t * out, int n)
{
  unsigned reg = 0x1ff; /* bit j = stage j */
  for (int i = 0; i < n; i++)
  {
    const unsigned b = ((reg >> 8) ^ (reg >> 4)) & 1u;
    reg = ((reg << 1) | b) & 0x1ff;
    out[i] = (uint8_t)b;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Puncturing vectors PI_1..PI_24 (protection/protTables.cpp:36-62, EN 300 401 table 13), bit j of
 * the word = V_PI[j].
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t PI_MASK[24] = {
  0x11111113u, 0x11131113u, 0x11131313u, 0x13131313u, 0x13131333u, 0x13331333u, 0x13333333u, 0x33333333u,
  0x33333337u, 0x33373337u, 0x33373737u, 0x37373737u, 0x37373777u, 0x37773777u, 0x37777777u, 0x77777777u,
  0x7777777fu, 0x777f777fu, 0x777f7f7fu, 0x7f7f7f7fu, 0x7f7f7fffu, 0x7fff7fffu, 0x7fffffffu, 0xffffffffu };

static inline int pi_bit(int pi, int j) { return (int)((PI_MASK[pi - 1] >> (j & 31)) & 1u); }

/* Appends the positions kept by `blocks` 128-bit blocks punctured with PI_pi; returns the new count. */
static int add_blocks(int32_t * addr, int n, int * pos, int blocks, int pi)
{
  for (int b = 0; b < blocks; b++)
    for (int j = 0; j < 128; j++)
    {
      if (pi_bit(pi, j)) addr[n++] = *pos;
      (*pos)++;
    }
  return n;
}
static int add_tail(int32_t * addr, int n, int * pos)
{
  for (int j = 0; j < 24; j++) /* PI_X = first 24 entries of PI_8 */
  {
    if (pi_bit(8, j)) addr[n++] = *pos;
    (*pos)++;
  }
  return n;
}

/* FIC: 21 blocks PI_16, 3 blocks PI_15, tail (decoder/fic_decoder.cpp:79-124). */
int dabo_fic_addresses(int32_t * addr, int cap)
{
  int32_t tmp[3096];
  int pos = 0, n = 0;
  n = add_blocks(tmp, n, &pos, 21, 16);
  n = add_blocks(tmp, n, &pos, 3, 15);
  n = add_tail(tmp, n, &pos);
  for (int i = 0; i < n && i < cap; i++) addr[i] = tmp[i];
  return n;
}

/* UEP profiles (protection/uep_protection.cpp:52-132, EN 300 401 table 15): bitrate, level, L1..4, PI1..4 */
typedef struct { int16_t bitrate, level, L[4], PI[4]; } uep_row;
static const uep_row UEP_TABLE[] = {
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
#define N_UEP ((int)(sizeof(UEP_TABLE) / sizeof(UEP_TABLE[0])))

/* Destination addresses of the kept bits inside the 4*24*bitRate+24 Viterbi block.
 * EEP: protection/eep_protection.cpp:43-151; UEP: protection/uep_protection.cpp:155-196. */
int dabo_protection_addresses(int short_form, int bit_rate, int prot_level, int32_t * addr, int cap)
{
  const int total = 4 * 24 * bit_rate + 24;
  int32_t * tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total + 4 * 128));
  int pos = 0, n = 0;
  if (short_form)
  {
    int idx = -1;
    for (int i = 0; i < N_UEP; i++) if (UEP_TABLE[i].bitrate == bit_rate && UEP_TABLE[i].level == prot_level) { idx = i; break; }
    if (idx < 0) idx = 1; /* reference falls back to row 1 (uep_protection.cpp:163-167) */
    for (int s = 0; s < 4; s++)
      if (UEP_TABLE[idx].L[s] > 0 && UEP_TABLE[idx].PI[s] > 0) n = add_blocks(tmp, n, &pos, UEP_TABLE[idx].L[s], UEP_TABLE[idx].PI[s]);
  }
  else
  {
    const int level = prot_level & 3;
    int L1 = 0, L2 = 0, P1 = 0, P2 = 0;
    if ((prot_level & 4) == 0) /* A profiles, n = bitRate/8 */
    {
      const int nn = bit_rate / 8;
      switch (level)
      {
      case 0: L1 = 6 * nn - 3; L2 = 3; P1 = 24; P2 = 23; break;
      case 1: if (nn == 1) { L1 = 5; L2 = 1; P1 = 13; P2 = 12; } else { L1 = 2 * nn - 3; L2 = 4 * nn + 3; P1 = 14; P2 = 13; } break;
      case 2: L1 = 6 * nn - 3; L2 = 3; P1 = 8; P2 = 7; break;
      default: L1 = 4 * nn - 3; L2 = 2 * nn + 3; P1 = 3; P2 = 2; break;
      }
    }
    else /* B profiles, n = bitRate/32 */
    {
      const int nn = bit_rate / 32;
      static const int8_t pb[4][2] = { {10, 9}, {6, 5}, {4, 3}, {2, 1} };
      L1 = 24 * nn - 3; L2 = 3; P1 = pb[level][0]; P2 = pb[level][1];
    }
    n = add_blocks(tmp, n, &pos, L1, P1);
    n = add_blocks(tmp, n, &pos, L2, P2);
  }
  n = add_tail(tmp, n, &pos);
  for (int i = 0; i < n && i < cap; i++) addr[i] = tmp[i];
  free(tmp);
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Viterbi K=7 rate 1/4, polynomials {109,79,83,109} (support/viterbi_spiral/viterbi_spiral.cpp:95-126,
 * viterbi_scalar.h:9-94): 32-bit metrics, start 0/1000, no renormalisation, path 0 kept on ties,
 * chain back from state 0.
 * ---------------------------------------------------------------------------------------------- */
static const int VIT_POLY[4] = { 109, 79, 83, 109 };
static uint8_t g_branch[4][32]; /* expected symbol (0/255) of butterfly i for the upper-left branch */
static int g_branch_ready = 0;

static int parity32(unsigned x) { x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return (int)(x & 1u); }

static void branch_init(void)
{
  if (g_branch_ready) return;
  /* Butterfly i joins old states i and i+32 to new states 2i, 2i+1. The branch old i -> new 2i has
   * encoder register (2i) with input 0, i.e. bits of (2i) & poly decide the expected output. This
   * reproduces Branchtable of viterbi_spiral.cpp:27-37. */
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 32; i++) g_branch[j][i] = parity32((unsigned)((2 * i) & VIT_POLY[j])) ? 255 : 0;
  g_branch_ready = 1;
}

void dabo_viterbi(const int16_t * in, int frame_bits, uint8_t * out)
{
  const int steps = frame_bits + 6;
  uint64_t * dec = (uint64_t *)calloc((size_t)steps, sizeof(uint64_t));
  int32_t ma[64], mb[64];
  int32_t * oldm = ma, * newm = mb;
  branch_init();
  for (int i = 0; i < 64; i++) oldm[i] = 1000;
  oldm[0] = 0;
  for (int t = 0; t < steps; t++)
  {
    int sym[4];
    for (int j = 0; j < 4; j++)
    {
      int v = (int16_t)(in[4 * t + j] + 127); /* `i16 tmp; tmp += 127;` wraps in 16 bits (viterbi_scalar.h:34-40) */
      sym[j] = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
    uint64_t d = 0;
    for (int i = 0; i < 32; i++)
    {
      int metric = 0;
      for (int j = 0; j < 4; j++) metric += g_branch[j][i] ^ sym[j];
      const int anti = 1020 - metric;
      const int32_t m0 = oldm[i] + metric, m1 = oldm[i + 32] + anti;
      const int32_t m2 = oldm[i] + anti, m3 = oldm[i + 32] + metric;
      const int d0 = (m0 - m1) > 0, d1 = (m2 - m3) > 0;
      newm[2 * i] = d0 ? m1 : m0;
      newm[2 * i + 1] = d1 ? m3 : m2;
      d |= (uint64_t)(d0 | (d1 << 1)) << (2 * i);
    }
    dec[t] = d;
    int32_t * tmp = oldm; oldm = newm; newm = tmp;
  }
  unsigned state = 0; /* 8 bit window, state index = window >> 2 */
  for (int t = frame_bits - 1; t >= 0; t--)
  {
    const unsigned k = (unsigned)((dec[t + 6] >> (state >> 2)) & 1u);
    state = (state >> 1) | (k << 7);
    out[t] = (uint8_t)k;
  }
  free(dec);
}

/* Re-encode and compare with the signs of the unpunctured inputs (viterbi_spiral.cpp:128-164). */
void dabo_viterbi_ber(const int16_t * in, const uint8_t * punct, const uint8_t * out_bits, int frame_bits, int * bits, int * errors)
{
  unsigned sr = 0;
  for (int i = 0; i < frame_bits + 6; i++)
  {
    sr = ((sr << 1) | (i < frame_bits ? out_bits[i] : 0u)) & 0xff;
    for (int j = 0; j < 4; j++)
    {
      if (!punct[4 * i + j]) continue;
      (*bits)++;
      if ((in[4 * i + j] > 0) != parity32(sr & (unsigned)VIT_POLY[j])) (*errors)++;
    }
  }
}

/* Protection::deconvolve (protection/protection.cpp:46-59): scatter, punctured positions stay 0. */
typedef struct
{
  int bit_rate, n_addr, block_len;
  int32_t * addr;
  int16_t * block;
} protection_t;

static protection_t * protection_new(int short_form, int bit_rate, int prot_level)
{
  protection_t * p = (protection_t *)calloc(1, sizeof(*p));
  p->bit_rate = bit_rate;
  p->block_len = 4 * 24 * bit_rate + 24;
  p->addr = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->block_len);
  p->n_addr = dabo_protection_addresses(short_form, bit_rate, prot_level, p->addr, p->block_len);
  p->block = (int16_t *)calloc((size_t)p->block_len, sizeof(int16_t));
  return p;
}
static void protection_free(protection_t * p) { free(p->addr); free(p->block); free(p); }
static void protection_run(protection_t * p, const int16_t * in, uint8_t * out)
{
  for (int i = 0; i < p->n_addr; i++) p->block[p->addr[i]] = in[i];
  dabo_viterbi(p->block, 24 * p->bit_rate, out);
}

int dabo_protection(int short_form, int bit_rate, int prot_level, const int16_t * in, int in_len, uint8_t * out)
{
  (void)in_len;
  protection_t * p = protection_new(short_form, bit_rate, prot_level);
  protection_run(p, in, out);
  protection_free(p);
  return 1;
}

/* ------------------------------------------------------------------------------------------------
 * CRC (backend/crc.cpp:75-132): CCITT x^16+x^12+x^5+1, all-ones preset, complemented remainder.
 * ---------------------------------------------------------------------------------------------- */
uint16_t dabo_calc_crc(const uint8_t * bytes, int n)
{
  unsigned crc = 0xffff;
  for (int i = 0; i < n; i++)
  {
    crc ^= (unsigned)bytes[i] << 8;
    for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xffff : (crc << 1) & 0xffff;
  }
  return (uint16_t)(~crc & 0xffff);
}

/* Bit-serial check over a message whose last 16 bits are the complemented CRC (crc.cpp:98-132). */
int dabo_check_crc_bits(const uint8_t * bits, int n)
{
  unsigned reg = 0xffff;
  for (int i = 0; i < n; i++)
  {
    const unsigned inb = (unsigned)(bits[i] ^ (i >= n - 16 ? 1 : 0)) & 1u;
    const unsigned fb = ((reg >> 15) & 1u) ^ inb;
    reg = (reg << 1) & 0xffff;
    if (fb) reg ^= 0x1021;
  }
  return reg == 0;
}

/* ------------------------------------------------------------------------------------------------
 * FIC (decoder/fic_decoder.cpp:143-262).
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  int32_t addr[FIC_VIT_IN];
  uint8_t punct[3096];
  uint8_t prbs[FIC_VIT_OUT];
  int16_t in[FIC_VIT_IN];
  int16_t block[3096];
  uint8_t bits[4 * FIC_VIT_OUT];
  uint8_t valid[4];
  int index, fic_idx, ratio, ber_bits, ber_errors, block_count, good_fibs;
} fic_t;

void * dabo_fic_new(void)
{
  fic_t * f = (fic_t *)calloc(1, sizeof(*f));
  dabo_fic_addresses(f->addr, FIC_VIT_IN);
  for (int i = 0; i < FIC_VIT_IN; i++) f->punct[f->addr[i]] = 1;
  dabo_prbs(f->prbs, FIC_VIT_OUT);
  return f;
}
void dabo_fic_free(void * h) { free(h); }

static void fic_one(fic_t * f)
{
  uint8_t * out = &f->bits[f->fic_idx * FIC_VIT_OUT];
  for (int i = 0; i < FIC_VIT_IN; i++) f->block[f->addr[i]] = f->in[i];
  dabo_viterbi(f->block, FIC_VIT_OUT, out);
  dabo_viterbi_ber(f->block, f->punct, out, FIC_VIT_OUT, &f->ber_bits, &f->ber_errors);
  if (++f->block_count == 40) { f->block_count = 0; f->ber_errors /= 2; f->ber_bits /= 2; }
  for (int i = 0; i < FIC_VIT_OUT; i++) out[i] ^= f->prbs[i];
  f->valid[f->fic_idx] = 1;
  for (int fib = 0; fib < 3; fib++)
  {
    if (dabo_check_crc_bits(&out[fib * FIB_BITS], FIB_BITS)) { f->good_fibs++; if (f->ratio < 10) f->ratio++; }
    else { f->valid[f->fic_idx] = 0; if (f->ratio > 0) f->ratio--; }
  }
}

void dabo_fic_process_block(void * h, const int16_t * soft, int sym_idx)
{
  fic_t * f = (fic_t *)h;
  if (sym_idx == 1) { f->index = 0; f->fic_idx = 0; }
  for (int i = 0; i < BITS_PER_SYM; i++)
  {
    f->in[f->index++] = soft[i];
    if (f->index >= FIC_VIT_IN) { fic_one(f); f->index = 0; f->fic_idx++; }
  }
}

void dabo_fic_get(void * h, uint8_t bits[3072], uint8_t valid[4], int * ratio_percent, int * ber_bits, int * ber_errors)
{
  fic_t * f = (fic_t *)h;
  memcpy(bits, f->bits, 3072);
  memcpy(valid, f->valid, 4);
  *ratio_percent = f->ratio * 10;
  *ber_bits = f->ber_bits;
  *ber_errors = f->ber_errors;
}

/* ------------------------------------------------------------------------------------------------
 * MSC backend (backend/backend.cpp:39-161): 16-deep time de-interleaver, EEP/UEP, dispersal.
 * ---------------------------------------------------------------------------------------------- */
static const int TIME_MAP[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };

typedef struct
{
  int sub_ch_id, start_cu, size_cu, fragment, bit_rate, count, widx;
  int16_t * mem[16];
  int16_t * tmp;
  uint8_t * prbs;
  protection_t * prot;
} backend_t;

void * dabo_backend_new(int sub_ch_id, int start_cu, int size_cu, int short_form, int prot_level, int bit_rate)
{
  backend_t * b = (backend_t *)calloc(1, sizeof(*b));
  b->sub_ch_id = sub_ch_id; b->start_cu = start_cu; b->size_cu = size_cu; b->bit_rate = bit_rate;
  b->fragment = size_cu * 64;
  for (int i = 0; i < 16; i++) b->mem[i] = (int16_t *)calloc((size_t)b->fragment, sizeof(int16_t));
  b->tmp = (int16_t *)calloc((size_t)b->fragment, sizeof(int16_t));
  b->prbs = (uint8_t *)malloc((size_t)(24 * bit_rate));
  dabo_prbs(b->prbs, 24 * bit_rate);
  b->prot = protection_new(short_form, bit_rate, prot_level);
  return b;
}
void dabo_backend_free(void * h)
{
  backend_t * b = (backend_t *)h;
  for (int i = 0; i < 16; i++) free(b->mem[i]);
  free(b->tmp); free(b->prbs); protection_free(b->prot); free(b);
}
int dabo_backend_process(void * h, const int16_t * frag, uint8_t * out)
{
  backend_t * b = (backend_t *)h;
  for (int i = 0; i < b->fragment; i++)
  {
    b->tmp[i] = b->mem[(b->widx + TIME_MAP[i & 15]) & 15][i];
    b->mem[b->widx][i] = frag[i];
  }
  b->widx = (b->widx + 1) & 15;
  if (b->count <= 15) { b->count++; return 0; }
  protection_run(b->prot, b->tmp, out);
  for (int i = 0; i < 24 * b->bit_rate; i++) out[i] ^= b->prbs[i];
  return 1;
}

/* ------------------------------------------------------------------------------------------------
 * OFDM decoder (ofdm/ofdm_decoder.cpp:90-355), scalar default build.
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  int soft_type;
  int16_t bin[K_CARR];
  cf32 ref[T_U];
  float integ[K_CARR], stddev[K_CARR], mean_pow[K_CARR], mean_sigma[K_CARR], null_pow[T_U];
  float mean_value, mean_pow_all;
  /* LCD statistics (ofdm_decoder.cpp:154-158,296-300,326-352; members ofdm_decoder.h:86-88,103) */
  int show_cnt, next_shown, n_lcd;
  float sigma_sq_freq_corr;
  float lcd[6]; /* last SLcdData: CurOfdmSymbolNo, MeanSigmaSqFreqCorr, SNR, MER, TestData1, TestData2 */
} ofdm_t;

void dabo_ofdm_reset(void * h)
{
  ofdm_t * d = (ofdm_t *)h;
  memset(d->integ, 0, sizeof(d->integ)); memset(d->stddev, 0, sizeof(d->stddev));
  memset(d->mean_pow, 0, sizeof(d->mean_pow)); memset(d->mean_sigma, 0, sizeof(d->mean_sigma));
  memset(d->null_pow, 0, sizeof(d->null_pow));
  d->mean_pow_all = 1.0f; /* mMeanValue is NOT reset (ofdm_decoder.cpp:90-101) */
}
void * dabo_ofdm_new(int soft_bit_type)
{
  ofdm_t * d = (ofdm_t *)calloc(1, sizeof(*d));
  d->soft_type = soft_bit_type;
  d->mean_value = 1.0f;
  d->next_shown = 1;
  dabo_freq_interleaver(d->bin);
  dabo_ofdm_reset(d);
  return d;
}
void dabo_ofdm_free(void * h) { free(h); }
void dabo_ofdm_store_reference_symbol_0(void * h, const float * fft) { memcpy(((ofdm_t *)h)->ref, fft, sizeof(cf32) * T_U); }

void dabo_ofdm_store_null_symbol_without_tii(void * h, const float * fft)
{
  ofdm_t * d = (ofdm_t *)h;
  const cf32 * x = (const cf32 *)fft;
  const float min_pow = (1.0f / 32767.0f) * (1.0f / 32767.0f);
  for (int idx = -K_CARR / 2; idx < K_CARR / 2; idx++)
  {
    const int b = idx < 0 ? idx + T_U : idx + 1;
    const float p = x[b].re * x[b].re + x[b].im * x[b].im + min_pow;
    d->null_pow[b] += 0.05f * (p - d->null_pow[b]);
  }
}

/* The figures signal_show_lcd_data carries, from the CURRENT state (ofdm_decoder.cpp:326-345, _compute_noise_Power :357-371):
 * out = { MER dB, SNR dB, mMeanValue, mMeanPowerOvrAll, noise power, sqrt(mMeanSigmaSqFreqCorr) } */
void dabo_ofdm_quality(void * h, float out[6])
{
  const ofdm_t * d = (const ofdm_t *)h;
  float sum_noise = 0.0f;
  for (int idx = -K_CARR / 2; idx < K_CARR / 2; idx++) sum_noise += d->null_pow[idx < 0 ? idx + T_U : idx + 1];
  if (sum_noise == 0.0f) sum_noise = (1.0f / 32767.0f) * (1.0f / 32767.0f) * (float)K_CARR;
  const float noise = sum_noise / (float)K_CARR;
  float snr = (d->mean_pow_all - noise) / noise;
  if (snr <= 0.0f) snr = 0.1f;
  float sd = 0.0f;
  for (int k = 0; k < K_CARR; k++) sd += d->stddev[k];
  sd /= (float)K_CARR;
  out[0] = 10.0f * log10f(PI_4_F * PI_4_F / sd);
  out[1] = 10.0f * log10f(snr);
  out[2] = d->mean_value;
  out[3] = d->mean_pow_all;
  out[4] = noise;
  out[5] = sqrtf(d->sigma_sq_freq_corr);
}

/* arg folded into [0, pi/2) (common/glob_defs.h:173-182) */
static float first_quadrant(float ph)
{
  if (ph < 0.0f) ph += PI_F;
  return fmodf(ph, PI_2_F);
}

void dabo_ofdm_decode_symbol(void * h, const float * fft, int sym_idx, float phase_corr, float clock_err, int16_t * out)
{
  ofdm_t * d = (ofdm_t *)h;
  const cf32 * x = (const cf32 *)fft;
  const float alpha = 0.005f;
  float sum = 0.0f;
  d->show_cnt++;
  const int show_stat = d->show_cnt > 5 * 76 && sym_idx == d->next_shown; /* ofdm_decoder.cpp:157 */
  for (int k = 0; k < K_CARR; k++)
  {
    int b = d->bin[k], rel = b;
    if (b < 0) { rel += K_CARR / 2; b += T_U; } else rel += K_CARR / 2 - 1;
    const float ref_abs = cabs32(d->ref[b]);
    cf32 raw = cmul(x[b], cconj(d->ref[b]));
    raw.re /= ref_abs; raw.im /= ref_abs;

    const float perr = clock_err / 1024.0f * PI_F * (float)(K_CARR / 2 - rel) / (float)(K_CARR / 2) + d->integ[k];
    /* cmplx_from_phase2(-perr): minimax sin/cos, deliberately not unit modulus (ofdm_decoder.cpp:70-88) */
    const float a = -perr, a2 = a * a;
    const cf32 rot = { 0.9994032382965087890625f + a2 * (a2 * 3.679168224334716796875e-2f + -0.495580852031707763671875f),
                       a * (a2 * -0.16034401953220367431640625f + 0.99903142452239990234375f) };
    const cf32 z = cmul(raw, rot);

    const float ph = first_quadrant(atan2f(z.im, z.re));
    d->integ[k] += 0.2f * alpha * (ph - PI_4_F);
    d->integ[k] = clampf_sym(d->integ[k], RAD_PER_DEG_F * 20.0f);
    const float dv = ph - PI_4_F;
    d->stddev[k] += alpha * (dv * dv - d->stddev[k]);
    const float pw = z.re * z.re + z.im * z.im;
    d->mean_pow[k] += alpha * (pw - d->mean_pow[k]);
    d->mean_pow_all += alpha / (float)K_CARR * (pw - d->mean_pow_all);

    const float lvl = sqrtf(d->mean_pow[k]);
    const float axis = lvl * 0.70710678118654752440084436210485f;
    const float dr = fabsf(z.re) - axis, di = fabsf(z.im) - axis;
    d->mean_sigma[k] += alpha * (dr * dr + di * di - d->mean_sigma[k]);
    float sig = d->mean_pow[k] - d->null_pow[b];
    if (sig <= 0.0f) sig = 0.1f;

    cf32 r;
    float w2;
    if (d->soft_type == 2) { r.re = z.re * ref_abs; r.im = z.im * ref_abs; w2 = -140.0f / d->mean_value; }
    else if (d->soft_type == 1)
    {
      float w1 = ref_abs / d->mean_sigma[k];
      w1 /= d->null_pow[b] / sig + 0.7f;
      r.re = z.re * w1; r.im = z.im * w1; w2 = -140.0f / d->mean_value;
    }
    else
    {
      const float zabs = sqrtf(pw);
      float w1 = sqrtf(zabs * ref_abs) * lvl;
      w1 /= d->null_pow[b] / sig + 0.7f;
      w1 /= d->mean_sigma[k] * zabs;
      r.re = z.re * w1; r.im = z.im * w1; w2 = -100.0f / d->mean_value;
    }
    out[k] = (int16_t)(r.re * w2);
    out[K_CARR + k] = (int16_t)(r.im * w2);
    sum += cabs32(r);
  }
  d->mean_value = sum / (float)K_CARR;
  if (sym_idx == 1)
  {
    const float fc = phase_corr / (2.0f * PI_F) * 1000.0f; /* cCarrDiff = 1000 Hz */
    d->sigma_sq_freq_corr += 0.2f * (fc * fc - d->sigma_sq_freq_corr);
  }
  if (show_stat)
  {
    float q[6];
    dabo_ofdm_quality(d, q);
    d->lcd[0] = (float)(sym_idx + 1); d->lcd[1] = q[5]; d->lcd[2] = q[1]; d->lcd[3] = q[0];
    d->lcd[4] = d->mean_value; d->lcd[5] = phase_corr / (2.0f * PI_F) * 1000.0f;
    d->n_lcd++;
    d->show_cnt = 0;
    d->next_shown = (d->next_shown + 1) % 76;
    if (d->next_shown == 0) d->next_shown = 1;
  }
  memcpy(d->ref, x, sizeof(cf32) * T_U);
}

/* last emitted SLcdData; returns the number of emissions so far */
int dabo_ofdm_lcd(void * h, float out[6])
{
  ofdm_t * d = (ofdm_t *)h;
  memcpy(out, d->lcd, sizeof(d->lcd));
  return d->n_lcd;
}

void dabo_ofdm_get_state(void * h, int which, float * out)
{
  ofdm_t * d = (ofdm_t *)h;
  switch (which)
  {
  case 0: memcpy(out, d->integ, sizeof(d->integ)); break;
  case 1: memcpy(out, d->stddev, sizeof(d->stddev)); break;
  case 2: memcpy(out, d->mean_pow, sizeof(d->mean_pow)); break;
  case 3: memcpy(out, d->mean_sigma, sizeof(d->mean_sigma)); break;
  case 4: memcpy(out, d->null_pow, sizeof(d->null_pow)); break;
  case 5: out[0] = d->mean_value; out[1] = d->mean_pow_all; break;
  default: break;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Phase reference correlation and coarse AFC (ofdm/phasereference.cpp:58-66,87-280).
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  cf32 prs[T_U];
  cf32 ref_arg_conj[T_U];
} phaseref_t;

static void relative_phase(cf32 * out, const cf32 * in)
{
  for (int i = 0; i < T_U - 1; i++) out[i] = cmul(cconj(in[i]), in[i + 1]);
  out[T_U - 1].re = out[T_U - 1].im = 0.0f;
}

void * dabo_phaseref_new(void)
{
  phaseref_t * p = (phaseref_t *)calloc(1, sizeof(*p));
  cf32 a[T_U], b[T_U];
  build_phase_table(p->prs);
  relative_phase(a, p->prs);
  fft2048(a, b, +1);
  for (int i = 0; i < T_U; i++) p->ref_arg_conj[i] = cconj(b[i]);
  return p;
}
void dabo_phaseref_free(void * h) { free(h); }

int dabo_phaseref_correlate(void * h, const float * samples, float threshold, int strongest_peak)
{
  phaseref_t * p = (phaseref_t *)h;
  static cf32 a[T_U], b[T_U];
  static float mag[T_U];
  fft2048((const cf32 *)samples, a, -1);
  for (int i = 0; i < T_U; i++) b[i] = cmul(a[i], cconj(p->prs[i]));
  fft2048(b, a, +1);
  float sum = 0.0f;
  for (int i = 0; i < T_U; i++) { mag[i] = cabs32(a[i]); sum += mag[i]; }
  sum /= (float)T_U;
  if (sum == 0.0f) return -1;
  const int lo = T_G - 250, hi = T_G + 500;
  int first = -1, best = -1;
  float best_v = -1000.0f;
  for (int i = lo; i < hi; i++)
  {
    if (!(mag[i] / sum > threshold)) continue;
    int is_peak = 1;
    for (int j = 1; j < 10 && i + j < hi; j++) if (mag[i + j] > mag[i]) { is_peak = 0; break; }
    if (!is_peak) continue;
    if (first < 0) first = i;
    if (mag[i] > best_v) { best_v = mag[i]; best = i; }
    i += 10;
  }
  if (best_v / sum < threshold) return -1;
  return strongest_peak ? best : first;
}

int dabo_phaseref_estimate_offset(void * h, const float * fft)
{
  phaseref_t * p = (phaseref_t *)h;
  static cf32 a[T_U], b[T_U];
  relative_phase(a, (const cf32 *)fft);
  fft2048(a, b, +1);
  for (int i = 0; i < T_U; i++) a[i] = cmul(b[i], p->ref_arg_conj[i]);
  fft2048(a, b, -1);
  int index = 100000;
  float mx = 0.0f, avg = 0.0f;
  for (int i = -70; i <= 70; i++)
  {
    const float v = cabs32(b[(T_U + i) % T_U]);
    if (v > mx) { mx = v; index = i; }
    avg += v;
  }
  avg /= 141.0f;
  if (mx < avg * 5.0f) return 100000;
  float pk[3], pk_sum = 0.0f;
  for (int i = 0; i < 3; i++) { pk[i] = cabs32(b[(T_U + index + i - 1) % T_U]); pk_sum += pk[i]; }
  const float offset = (float)index + (pk[2] - pk[0]) / pk_sum;
  return (int32_t)(offset * (float)CARR_DIFF);
}

/* ------------------------------------------------------------------------------------------------
 * Whole chain: SampleReader (ofdm/sample_reader.cpp:44-50,102-297), TimeSyncer (ofdm/timesyncer.cpp:40-90),
 * DabProcessor (main/dab_processor.cpp:110-442), MscHandler (backend/msc_handler.cpp:148-168).
 * ---------------------------------------------------------------------------------------------- */
/* The part of FibDecoder::process_FIB that get_cif_count() depends on (fib_decoder.cpp:59-106 FIG walk,
 * fib_decoder_fig0.cpp:89-113): the CIF counter of the last FIG 0/0 among the CRC-good FIBs of one frame's FIC. */
static void scan_cif_count(const uint8_t * fib_bits, int * cif_hi, int * cif_lo)
{
  for (int k = 0; k < 12; k++)
  {
    const uint8_t * fib = fib_bits + (k / 3) * 768 + (k % 3) * 256;
    if (!dabo_check_crc_bits(fib, 256)) continue;
    int done = 0;
    while (done < 30)
    {
      const uint8_t * d = fib + 8 * done;
      unsigned type = 0, len = 0, ext = 0, hi = 0, lo = 0;
      for (int b = 0; b < 3; b++) type = (type << 1) | d[b];
      for (int b = 3; b < 8; b++) len = (len << 1) | d[b];
      if ((type == 7 && len == 31) || done + (int)len + 1 > 30) break;
      for (int b = 11; b < 16; b++) ext = (ext << 1) | d[b];
      if (type == 0 && ext == 0 && len >= 5)
      {
        for (int b = 35; b < 40; b++) hi = (hi << 1) | d[b];
        for (int b = 40; b < 48; b++) lo = (lo << 1) | d[b];
        *cif_hi = (int)hi; *cif_lo = (int)lo;
      }
      done += (int)len + 1;
    }
  }
}

typedef struct
{
  dabo_frame_info info;
  uint8_t fib[3072];
  int16_t * soft;
  float * fft;
} frame_rec;

typedef struct
{
  int sub_ch_id, start_frame;
  backend_t * be;
  uint8_t * out;
  int64_t n, cap;
} msc_ch;

typedef struct
{
  dabo_chain_cfg cfg;
  const cf32 * iq;
  int64_t n_samples, pos, single_reads;
  cf32 * osc;
  int32_t phase;
  float s_level;
  int end_of_data;

  ofdm_t * ofdm;
  fic_t * fic;
  phaseref_t * pref;
  msc_ch * ch;
  int n_ch;
  int16_t cif[CIF_BITS];
  void * eti;                 /* eti_t when cfg.eti_path is set */

  frame_rec * frames;
  int n_frames, cap_frames;
  int dip_found, no_dip, sync_attempts;
  int cif_hi, cif_lo;         /* cfg.track_cif: what a real FIB decoder would hold (0, 0 until the first FIG 0/0) */
  double seconds;
} chain_t;

/* get_samples: level IIR on |v| then mixing with the integer-Hz oscillator table (sample_reader.cpp:212-282). */
static int read_samples(chain_t * c, cf32 * dst, int n, float freq_hz)
{
  if (c->pos + n > c->n_samples) { c->end_of_data = 1; return 0; }
  const int32_t f = (int32_t)roundf(freq_hz);
  for (int i = 0; i < n; i++)
  {
    const cf32 v = c->iq[c->pos + i];
    const float a = cabs32(v);
    c->s_level += 0.00001f * (a - c->s_level);
    c->phase -= f;
    c->phase = (c->phase + FS) % FS;
    dst[i] = cmul(v, c->osc[c->phase]);
  }
  c->pos += n;
  if (n == 1) c->single_reads++;
  return 1;
}

/* returns 1 established, 0 no dip, 2 no end of dip, -1 end of data (timesyncer.cpp:40-90) */
static int time_sync(chain_t * c)
{
  float env[4096];
  float level = 0.0f;
  int idx = 0;
  cf32 s;
  for (int i = 0; i < 50; i++)
  {
    if (!read_samples(c, &s, 1, 0.0f)) return -1;
    env[idx] = cabs32(s); level += env[idx]; idx++;
  }
  int counter = 0;
  while (level / 50.0f > 0.55f * c->s_level)
  {
    if (!read_samples(c, &s, 1, 0.0f)) return -1;
    env[idx] = cabs32(s);
    level += env[idx] - env[(idx - 50) & 4095];
    idx = (idx + 1) & 4095;
    if (++counter > T_FRAME) return 0;
  }
  counter = 0;
  while (level / 50.0f < 0.75f * c->s_level)
  {
    if (!read_samples(c, &s, 1, 0.0f)) return -1;
    env[idx] = cabs32(s);
    level += env[idx] - env[(idx - 50) & 4095];
    idx = (idx + 1) & 4095;
    if (++counter > T_NULL + 50 + 20) return 2;
  }
  return 1;
}

static frame_rec * new_frame(chain_t * c)
{
  if (c->n_frames == c->cap_frames)
  {
    c->cap_frames = c->cap_frames ? 2 * c->cap_frames : 64;
    c->frames = (frame_rec *)realloc(c->frames, sizeof(frame_rec) * (size_t)c->cap_frames);
  }
  frame_rec * f = &c->frames[c->n_frames++];
  memset(f, 0, sizeof(*f));
  if (c->cfg.tap_soft_bits) f->soft = (int16_t *)calloc((size_t)75 * BITS_PER_SYM, sizeof(int16_t));
  if (c->cfg.tap_fft) f->fft = (float *)calloc((size_t)77 * T_U * 2, sizeof(float));
  return f;
}

/* ------------------------------------------------------------------------------------------------
 * ETI-NI generator (eti_handler/eti_generator.cpp:63-204 process_block, :207-308 _init_eti, :330-412 _process_cif /
 * _process_sub_channel). Whole-CIF time de-interleaver with the same map as Backend, but emission starts one CIF
 * earlier and the start-up bookkeeping (`amount`, `Minor`, `index_Out`) drops the 16th CIF from the history; the CIF
 * counter comes from the FIB decoder (cif_hi / cif_lo of the configuration, the harness' stub returns 0, 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  int running, index_out, amount, minor, cif_hi, cif_lo;
  int16_t cif_in[CIF_BITS];
  int16_t (*cif_vec)[CIF_BITS];   /* 16 CIFs, zero at start (file-scope statics in the reference, eti_generator.cpp:9-24) */
  int16_t * temp;
  uint8_t fib_vec[16][96];
  protection_t * prot[64];
  uint8_t * desc[64];
  uint8_t * out;
  int64_t n, cap;
} eti_t;

static eti_t * eti_new(void)
{
  eti_t * e = (eti_t *)calloc(1, sizeof(*e));
  e->cif_vec = (int16_t (*)[CIF_BITS])calloc(16, sizeof(int16_t) * CIF_BITS);
  e->temp = (int16_t *)calloc(CIF_BITS, sizeof(int16_t));
  e->cif_hi = e->cif_lo = e->minor = -1;
  return e;
}

static void eti_free(eti_t * e)
{
  if (!e) return;
  for (int i = 0; i < 64; i++) { if (e->prot[i]) protection_free(e->prot[i]); free(e->desc[i]); }
  free(e->cif_vec); free(e->temp); free(e->out); free(e);
}

/* eti_generator.cpp:207-308 */
static int eti_init_frame(uint8_t * eti, int cif_hi, int cif_lo, int minor, const int32_t * subch, int n_subch)
{
  int fill = 0;
  cif_lo += minor;
  if (cif_lo >= 250) { cif_lo %= 250; cif_hi++; }
  if (cif_hi >= 20) cif_hi = 20;
  eti[fill++] = 0xFF;                                         /* ERR: error level 0 */
  if (cif_lo & 1) { eti[fill++] = 0xf8; eti[fill++] = 0xc5; eti[fill++] = 0x49; }
  else { eti[fill++] = 0x07; eti[fill++] = 0x3a; eti[fill++] = 0xb6; }
  eti[fill++] = (uint8_t)cif_lo;                              /* FCT */
  int nst = 0, fl = 0;
  for (int i = 0; i < n_subch; i++) { nst++; fl += (subch[7 * i + 5] * 3) / 4; }
  fl += nst + 1 + 24;                                         /* STC + EOH + FIC (mode I), in words */
  eti[fill++] = (uint8_t)((1 << 7) | nst);                    /* FICF | NST */
  const int fp = ((cif_hi * 250) + cif_lo) % 8;
  eti[fill++] = (uint8_t)((fp << 5) | (0x01 << 3) | ((fl & 0x700) >> 8));
  eti[fill++] = (uint8_t)(fl & 0xff);
  for (int i = 0; i < n_subch; i++)
  {
    const int32_t * sc = subch + 7 * i;
    const int scid = sc[0], sad = sc[1];
    const int tpl = sc[3] ? (0x10 | (sc[4] - 1)) : (0x20 | sc[4]);
    const int stl = sc[5] * 3 / 8;
    eti[fill++] = (uint8_t)((scid << 2) | ((sad & 0x300) >> 8));
    eti[fill++] = (uint8_t)(sad & 0xFF);
    eti[fill++] = (uint8_t)((tpl << 2) | ((stl & 0x300) >> 8));
    eti[fill++] = (uint8_t)(stl & 0xFF);
  }
  eti[fill++] = 0xFF; eti[fill++] = 0xFF;                     /* MNSC */
  const uint16_t hcrc = dabo_calc_crc(&eti[4], fill - 4);
  eti[fill++] = (uint8_t)(hcrc >> 8);
  eti[fill++] = (uint8_t)(hcrc & 0xff);
  return fill;
}

/* one OFDM symbol of soft bits, eti_generator.cpp:90-204 */
static void eti_block(eti_t * e, const dabo_chain_cfg * cfg, const uint8_t * fib_bits, const int16_t * soft, int sym)
{
  if (!e->running && sym == 1) e->running = 1;
  if (!e->running || sym < 4) return;
  if (sym == 4)
  {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 96; j++)
      {
        unsigned v = 0;
        for (int k = 0; k < 8; k++) v = (v << 1) | (fib_bits[i * 768 + 8 * j + k] & 1u);
        e->fib_vec[(e->index_out + i) & 15][j] = (uint8_t)v;
      }
    e->minor = 0;
    if (cfg->eti_cif_hi >= 0) { e->cif_hi = cfg->eti_cif_hi; e->cif_lo = cfg->eti_cif_lo; }
    else
    {
      /* a real FIB decoder behind get_cif_count(): the CIF counter of the last FIG 0/0 among the CRC-good FIBs so far
       * (fib_decoder.cpp:59-106 FIG walk, fib_decoder_fig0.cpp:89-113) */
      for (int k = 0; k < 12; k++)
      {
        const uint8_t * fib = fib_bits + (k / 3) * 768 + (k % 3) * 256;
        if (!dabo_check_crc_bits(fib, 256)) continue;
        int done = 0;
        while (done < 30)
        {
          const uint8_t * d = fib + 8 * done;
          unsigned type = 0, len = 0, ext = 0, hi = 0, lo = 0;
          for (int b = 0; b < 3; b++) type = (type << 1) | d[b];
          for (int b = 3; b < 8; b++) len = (len << 1) | d[b];
          if ((type == 7 && len == 31) || done + (int)len + 1 > 30) break;
          for (int b = 11; b < 16; b++) ext = (ext << 1) | d[b];
          if (type == 0 && ext == 0 && len >= 5)
          {
            for (int b = 35; b < 40; b++) hi = (hi << 1) | d[b];
            for (int b = 40; b < 48; b++) lo = (lo << 1) | d[b];
            e->cif_hi = (int)hi; e->cif_lo = (int)lo;
          }
          done += (int)len + 1;
        }
      }
    }
  }
  const int blk = (sym - 4) % 18;
  memcpy(&e->cif_in[blk * BITS_PER_SYM], soft, sizeof(int16_t) * BITS_PER_SYM);
  if (blk != 17) return;
  for (int i = 0; i < CIF_BITS; i++)
  {
    e->temp[i] = e->cif_vec[(e->index_out + TIME_MAP[i & 15]) & 15][i];
    e->cif_vec[e->index_out & 15][i] = e->cif_in[i];
  }
  if (e->amount < 15) { e->amount++; e->index_out = (e->index_out + 1) & 15; e->minor = -1; return; }
  if (e->cif_hi < 0 || e->cif_lo < 0) { e->minor = -1; return; }
  if (e->minor < 0) return;
  uint8_t v[6144];
  int off = eti_init_frame(v, e->cif_hi, e->cif_lo, e->minor, cfg->subch, cfg->n_subch);
  const int base = off;
  memcpy(&v[off], e->fib_vec[e->index_out], 96);
  off += 96;
  for (int i = 0; i < cfg->n_subch; i++)
  {
    const int32_t * sc = cfg->subch + 7 * i;
    const int id = sc[0] & 63, nb = 24 * sc[5];
    if (!e->prot[id])
    {
      e->prot[id] = protection_new(sc[3], sc[5], sc[4]);
      e->desc[id] = (uint8_t *)malloc((size_t)nb);
      dabo_prbs(e->desc[id], nb);
    }
    uint8_t * bits = (uint8_t *)calloc((size_t)nb, 1);
    protection_run(e->prot[id], &e->temp[sc[1] * 64], bits);
    for (int j = 0; j < nb / 8; j++)
    {
      unsigned t = 0;
      for (int k = 0; k < 8; k++) t = (t << 1) | ((bits[8 * j + k] ^ e->desc[id][8 * j + k]) & 1u);
      v[off++] = (uint8_t)t;
    }
    free(bits);
  }
  const uint16_t crc = dabo_calc_crc(&v[base], off - base);
  v[off++] = (uint8_t)(crc >> 8); v[off++] = (uint8_t)(crc & 0xff);
  v[off++] = 0xFF; v[off++] = 0xFF;                           /* RFU */
  v[off++] = 0xFF; v[off++] = 0xFF; v[off++] = 0xFF; v[off++] = 0xFF; /* TIST unused */
  memset(&v[off], 0x55, (size_t)(6144 - off));
  if (e->n + 6144 > e->cap) { e->cap = e->cap ? 2 * e->cap : 6144 * 64; e->out = (uint8_t *)realloc(e->out, (size_t)e->cap); }
  memcpy(e->out + e->n, v, 6144);
  e->n += 6144;
  e->index_out = (e->index_out + 1) & 15;
  e->minor++;
}

static void msc_block(chain_t * c, const int16_t * soft, int sym, int frames_done)
{
  const int blk = (sym - 4) % 18;
  memcpy(&c->cif[blk * BITS_PER_SYM], soft, sizeof(int16_t) * BITS_PER_SYM);
  if (blk < 17) return;
  for (int i = 0; i < c->n_ch; i++)
  {
    msc_ch * m = &c->ch[i];
    if (frames_done < m->start_frame) continue;
    if (m->be == NULL)
    {
      const int32_t * s = c->cfg.subch + 7 * i;
      m->be = (backend_t *)dabo_backend_new(s[0], s[1], s[2], s[3], s[4], s[5]);
    }
    const int nb = 24 * m->be->bit_rate;
    if (m->n + nb > m->cap) { m->cap = m->cap ? 2 * m->cap : (int64_t)nb * 64; m->out = (uint8_t *)realloc(m->out, (size_t)m->cap); }
    if (dabo_backend_process(m->be, &c->cif[m->be->start_cu * 64], m->out + m->n)) m->n += nb;
  }
}

void * dabo_chain_run(const float * iq, int64_t n_samples, const dabo_chain_cfg * cfg)
{
  chain_t * c = (chain_t *)calloc(1, sizeof(*c));
  struct timespec t0, t1;
  c->cfg = *cfg;
  c->iq = (const cf32 *)iq;
  c->n_samples = n_samples;
  c->s_level = 0.1f;
  c->osc = (cf32 *)malloc(sizeof(cf32) * FS);
  for (int i = 0; i < FS; i++)
  {
    c->osc[i].re = (float)cos(2.0 * M_PI * (double)i / (double)FS);
    c->osc[i].im = (float)sin(2.0 * M_PI * (double)i / (double)FS);
  }
  c->ofdm = (ofdm_t *)dabo_ofdm_new(cfg->soft_bit_type);
  c->fic = (fic_t *)dabo_fic_new();
  if (cfg->eti_path != NULL) c->eti = eti_new();
  c->pref = (phaseref_t *)dabo_phaseref_new();
  c->n_ch = cfg->scan_mode ? 0 : cfg->n_subch;
  c->ch = (msc_ch *)calloc((size_t)(c->n_ch > 0 ? c->n_ch : 1), sizeof(msc_ch));
  for (int i = 0; i < c->n_ch; i++) { c->ch[i].sub_ch_id = cfg->subch[7 * i]; c->ch[i].start_frame = cfg->subch[7 * i + 6]; }

  static cf32 buf[T_NULL], fft_in[T_U], fft_out[T_U];
  static int16_t bits[BITS_PER_SYM];
  float f_sync = 0.0f, f_bb = 0.0f, clock_err = 0.0f, phase_cp = 0.0f, thr = 0.0f;
  int sample_count = 0, frames_done = 0;
  enum { WAIT_SYNC, EVAL_SYNC, REST } state = WAIT_SYNC;

  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < 20 && !c->end_of_data; i++) read_samples(c, buf, T_U, 0.0f);
  while (!c->end_of_data)
  {
    if (state == WAIT_SYNC)
    {
      dabo_ofdm_reset(c->ofdm);
      sample_count = 0;
      thr = cfg->threshold;
      const int r = time_sync(c);
      if (r < 0) break;
      /* _state_wait_for_time_sync_marker (dab_processor.cpp:416-442): signal_no_dip_sync_found goes out once per eight
         NO_DIP_FOUND results in a row; NO_END_OF_DIP_FOUND and a success restart the count */
      if (r == 1) { c->dip_found++; c->sync_attempts = 0; state = EVAL_SYNC; }
      else if (r == 2) c->sync_attempts = 0;
      else if (++c->sync_attempts >= 8) { c->sync_attempts = 0; c->no_dip++; }
      clock_err = 0.0f;
    }
    else if (state == EVAL_SYNC)
    {
      if (!read_samples(c, buf, T_U, f_bb)) break;
      const int start = dabo_phaseref_correlate(c->pref, (const float *)buf, thr, cfg->strongest_peak);
      if (start < 0) { state = WAIT_SYNC; continue; }
      memmove(buf, &buf[start], sizeof(cf32) * (size_t)(T_U - start));
      if (start > 0 && !read_samples(c, &buf[T_U - start], start, f_bb)) break;
      sample_count = start + T_U;
      frame_rec * f = new_frame(c);
      f->info.sym0_pos = c->pos - T_U;
      f->info.start_index = start;
      state = REST;
    }
    else
    {
      frame_rec * f = &c->frames[c->n_frames - 1];
      f->info.fbb_sym0 = f_bb;
      f->info.fic_ratio_before = c->fic->ratio * 10;
      fft2048(buf, fft_out, -1);
      if (f->fft) memcpy(f->fft, fft_out, sizeof(cf32) * T_U);
      dabo_ofdm_store_reference_symbol_0(c->ofdm, (const float *)fft_out);
      int correction = 0;
      if (c->fic->ratio * 10 < 30)
      {
        correction = dabo_phaseref_estimate_offset(c->pref, (const float *)fft_out);
        if (correction != 100000)
        {
          f_sync += (float)correction;
          if (fabsf(f_sync) > 35000.0f) f_sync = 0.0f;
        }
        if (correction != 0) clock_err = 0.0f;
        f_bb = f_sync;
      }
      f->info.fbb_data = f_bb;
      f->info.clock_err = clock_err;

      cf32 corr = { 0.0f, 0.0f };
      int cut = 0;
      for (int sym = 1; sym < L_SYMS; sym++)
      {
        if (!read_samples(c, buf, T_S, f_bb)) { cut = 1; break; }
        sample_count += T_S;
        for (int i = T_U; i < T_S; i++)
        {
          const cf32 p = cmul(buf[i], cconj(buf[i - T_U]));
          corr.re += p.re; corr.im += p.im;
        }
        memcpy(fft_in, &buf[T_G], sizeof(cf32) * T_U);
        fft2048(fft_in, fft_out, -1);
        if (f->fft) memcpy(&f->fft[(size_t)sym * T_U * 2], fft_out, sizeof(cf32) * T_U);
        dabo_ofdm_decode_symbol(c->ofdm, (const float *)fft_out, sym, phase_cp, clock_err, bits);
        if (f->soft) memcpy(&f->soft[(size_t)(sym - 1) * BITS_PER_SYM], bits, sizeof(bits));
        if (sym <= 3) dabo_fic_process_block(c->fic, bits, sym);
        else if (!cfg->scan_mode) msc_block(c, bits, sym, frames_done);
        if (c->eti) eti_block((eti_t *)c->eti, cfg, c->fic->bits, bits, sym); /* dab_processor.cpp:352-355 */
      }
      if (cut) { c->n_frames--; break; }
      phase_cp = clampf_sym(atan2f(corr.im, corr.re), 20.0f * RAD_PER_DEG_F);
      f_sync += phase_cp / TWO_PI_F * (float)CARR_DIFF;
      f_bb = f_sync;

      if (!read_samples(c, buf, T_NULL, f_bb)) { c->n_frames--; break; }
      sample_count += T_NULL;
      f->info.fbb_null = f_bb;
      f->info.fsync = f_sync;
      f->info.phase_cp = phase_cp;
      f->info.fic_ratio_after = c->fic->ratio * 10;
      memcpy(f->fib, c->fic->bits, 3072);
      memcpy(f->info.fic_valid, c->fic->valid, 4);
      memcpy(fft_in, &buf[T_G], sizeof(cf32) * T_U);
      fft2048(fft_in, fft_out, -1);
      if (f->fft) memcpy(&f->fft[(size_t)76 * T_U * 2], fft_out, sizeof(cf32) * T_U);
      /* dab_processor.cpp:273-300: TII null symbols ((cif count & 7) >= 4) do not update the null power; without a FIB decoder
       * behind get_cif_count() (the harness' default stub) the count is 0 and every null symbol is a plain one */
      if (cfg->track_cif) scan_cif_count(c->fic->bits, &c->cif_hi, &c->cif_lo);
      if (!(cfg->track_cif && (((c->cif_hi * 250 + c->cif_lo) & 7) >= 4)))
        dabo_ofdm_store_null_symbol_without_tii(c->ofdm, (const float *)fft_out);
      if (correction == 0)
      {
        const float ce = clampf_sym((float)FS * ((float)sample_count / (float)T_FRAME - 1.0f), 307.2f);
        clock_err += 0.1f * (ce - clock_err);
      }
      frames_done++;
      state = EVAL_SYNC;
      thr = 2.0f * cfg->threshold;
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  c->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (c->eti && cfg->eti_path[0] != 0)
  {
    FILE * f = fopen(cfg->eti_path, "wb");
    if (f) { fwrite(((eti_t *)c->eti)->out, 1, (size_t)((eti_t *)c->eti)->n, f); fclose(f); }
  }
  return c;
}

void dabo_chain_free(void * h)
{
  chain_t * c = (chain_t *)h;
  for (int i = 0; i < c->n_frames; i++) { free(c->frames[i].soft); free(c->frames[i].fft); }
  for (int i = 0; i < c->n_ch; i++) { if (c->ch[i].be) dabo_backend_free(c->ch[i].be); free(c->ch[i].out); }
  free(c->frames); free(c->ch); free(c->osc);
  dabo_ofdm_free(c->ofdm); dabo_fic_free(c->fic); dabo_phaseref_free(c->pref);
  eti_free((eti_t *)c->eti);
  free(c);
}
int64_t dabo_chain_eti_size(void * h) { chain_t * c = (chain_t *)h; return c->eti ? ((eti_t *)c->eti)->n : 0; }
int64_t dabo_chain_eti_copy(void * h, uint8_t * out, int64_t cap)
{
  chain_t * c = (chain_t *)h;
  if (!c->eti) return 0;
  const int64_t n = ((eti_t *)c->eti)->n < cap ? ((eti_t *)c->eti)->n : cap;
  memcpy(out, ((eti_t *)c->eti)->out, (size_t)n);
  return n;
}
int dabo_chain_n_frames(void * h) { return ((chain_t *)h)->n_frames; }
void dabo_chain_frame_info(void * h, int frame, dabo_frame_info * out) { *out = ((chain_t *)h)->frames[frame].info; }
void dabo_chain_fib_bits(void * h, int frame, uint8_t out[3072]) { memcpy(out, ((chain_t *)h)->frames[frame].fib, 3072); }
int dabo_chain_soft_bits(void * h, int frame, int16_t * out)
{
  frame_rec * f = &((chain_t *)h)->frames[frame];
  if (!f->soft) return 0;
  memcpy(out, f->soft, sizeof(int16_t) * 75 * BITS_PER_SYM);
  return 1;
}
int dabo_chain_fft(void * h, int frame, float * out)
{
  frame_rec * f = &((chain_t *)h)->frames[frame];
  if (!f->fft) return 0;
  memcpy(out, f->fft, sizeof(float) * 77 * T_U * 2);
  return 1;
}
int dabo_chain_n_good_fibs(void * h) { return ((chain_t *)h)->fic->good_fibs; }
static msc_ch * find_ch(chain_t * c, int id)
{
  for (int i = 0; i < c->n_ch; i++) if (c->ch[i].sub_ch_id == id) return &c->ch[i];
  return NULL;
}
int64_t dabo_chain_msc_size(void * h, int id) { msc_ch * m = find_ch((chain_t *)h, id); return m ? m->n : 0; }
int64_t dabo_chain_msc_copy(void * h, int id, uint8_t * out, int64_t cap)
{
  msc_ch * m = find_ch((chain_t *)h, id);
  if (!m) return 0;
  const int64_t n = m->n < cap ? m->n : cap;
  memcpy(out, m->out, (size_t)n);
  return n;
}
void dabo_chain_counters(void * h, int64_t out[8])
{
  chain_t * c = (chain_t *)h;
  memset(out, 0, sizeof(int64_t) * 8);
  out[0] = c->dip_found; out[1] = c->no_dip; out[2] = c->single_reads; out[3] = c->pos;
}
double dabo_chain_seconds(void * h) { return ((chain_t *)h)->seconds; }
void dabo_chain_quality(void * h, float out[6]) { dabo_ofdm_quality(((chain_t *)h)->ofdm, out); }
int dabo_chain_lcd(void * h, float out[6]) { return dabo_ofdm_lcd(((chain_t *)h)->ofdm, out); }
