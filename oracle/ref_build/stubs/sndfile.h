#pragma once
typedef struct SNDFILE_tag SNDFILE;
typedef long sf_count_t;
static inline sf_count_t sf_writef_short(SNDFILE *, const short *, sf_count_t n) { return n; }
