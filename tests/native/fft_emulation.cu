// CPU emulation of the CTA-wide 2048-point FFT (dabstar_b200/csrc/fft2048.cuh): runs the same
// __host__ __device__ phase functions thread by thread and compares with a direct double-precision DFT.
// Built with nvcc as a host-only program; no GPU needed. Prints the maximum relative error.
#include "fft2048.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace dab;

int main()
{
  std::vector<float2> w(2048), x(2048);
  for (int m = 0; m < 2048; m++) w[m] = make_float2((float)cos(2.0 * M_PI * m / 2048.0), (float)-sin(2.0 * M_PI * m / 2048.0));
  srand(7);
  for (auto & s : x) s = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
  static float2 regs[FFT_THREADS][16];
  static FftTwiddles tw[FFT_THREADS];
  std::vector<float2> smem(FFT_SMEM_F2), out(2048), tw2(FFT_TW2_F2);
  for (int t = 0; t < FFT_THREADS; t++)
  {
    fft_load_twiddles(tw[t], w.data(), tw2.data(), t);
    for (int n1 = 0; n1 < 16; n1++) regs[t][n1] = x[128 * n1 + t];
  }
  for (int t = 0; t < FFT_THREADS; t++) fft_stage1(regs[t], tw[t], smem.data(), t);
  for (int t = 0; t < FFT_THREADS; t++) fft_stage2_read(regs[t], smem.data(), t);
  for (int t = 0; t < FFT_THREADS; t++) fft_stage2(regs[t], tw[t], smem.data(), t);
  for (int t = 0; t < FFT_THREADS; t++) fft_stage3(regs[t], smem.data(), t);
  std::vector<int> hit(2048, 0);
  for (int t = 0; t < FFT_THREADS; t++)
    for (int h = 0; h < 2; h++)
      for (int j2 = 0; j2 < 8; j2++) { out[fft_out_index(t, h, j2)] = regs[t][8 * h + j2]; hit[fft_out_index(t, h, j2)]++; }
  std::vector<int> slot(FFT_NAT_F2, 0);
  for (int k = 0; k < 2048; k++) { if (fft_nat(k) >= FFT_NAT_F2 || slot[fft_nat(k)]++) { printf("FAIL: padded index collision at %d\n", k); return 1; } }
  for (int k = 0; k < 2048; k++) if (hit[k] != 1) { printf("FAIL: output index %d written %d times\n", k, hit[k]); return 1; }
  double max_err = 0, max_mag = 0;
  for (int k = 0; k < 2048; k++)
  {
    double re = 0, im = 0;
    for (int n = 0; n < 2048; n++)
    {
      const double a = -2.0 * M_PI * (double)((k * n) % 2048) / 2048.0;
      re += x[n].x * cos(a) - x[n].y * sin(a);
      im += x[n].x * sin(a) + x[n].y * cos(a);
    }
    max_err = fmax(max_err, hypot(out[k].x - re, out[k].y - im));
    max_mag = fmax(max_mag, hypot(re, im));
  }
  printf("max_abs_err %.3e max_mag %.3e rel %.3e\n", max_err, max_mag, max_err / max_mag);
  return max_err / max_mag < 2e-6 ? 0 : 1;
}
