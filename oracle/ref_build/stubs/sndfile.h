// libsndfile stand-in: a "file" is an array of float frames in memory (what sf_readf_float delivers for any PCM format).
#pragma once
#include <cstring>
typedef long sf_count_t;
typedef struct SNDFILE_tag
{
  const float * frames; // interleaved I, Q
  sf_count_t n_frames, pos;
} SNDFILE;
static inline sf_count_t sf_writef_short(SNDFILE *, const short *, sf_count_t n) { return n; }
static inline sf_count_t sf_seek(SNDFILE * f, sf_count_t off, int whence)
{
  f->pos = whence == SEEK_SET ? off : (whence == SEEK_CUR ? f->pos + off : f->n_frames + off);
  if (f->pos < 0) f->pos = 0;
  if (f->pos > f->n_frames) f->pos = f->n_frames;
  return f->pos;
}
static inline sf_count_t sf_readf_float(SNDFILE * f, float * out, sf_count_t n)
{
  const sf_count_t m = n < f->n_frames - f->pos ? n : f->n_frames - f->pos;
  if (m > 0) memcpy(out, f->frames + 2 * f->pos, sizeof(float) * 2 * (size_t)m);
  f->pos += m > 0 ? m : 0;
  return m > 0 ? m : 0;
}
