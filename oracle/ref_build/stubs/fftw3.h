// FFTW3f is not vendored by the reference and not installed here: this declares the three calls
// the hot path uses; oracle/ref_build/link_stubs.cpp implements them (double precision radix-2).
#pragma once
typedef float fftwf_complex[2];
struct fftwf_plan_s;
typedef fftwf_plan_s * fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
extern "C" {
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex * in, fftwf_complex * out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
}
