// checks host_fft_epilogue_layout: slots unique, every write / read group conflict free
#include "tables.h"
#include "fft2048.cuh"
#include <cstdio>
#include <set>
#include <vector>
using namespace dab;
int main()
{
  std::vector<int16_t> bs(K_CARR), bin(K_CARR);
  host_freq_interleaver(bs.data());
  for (int k = 0; k < K_CARR; k++) bin[k] = (int16_t)(bs[k] < 0 ? bs[k] + T_U : bs[k]);
  std::vector<uint16_t> sw(16 * FFT_THREADS), sr(K_CARR);
  host_fft_epilogue_layout(bin.data(), sw.data(), sr.data());
  std::set<int> all(sw.begin(), sw.end());
  if (all.size() != 2048) { printf("write slots not unique: %zu\n", all.size()); return 1; }
  std::vector<int> carrier_of_bin(T_U, -1);
  for (int k = 0; k < K_CARR; k++) carrier_of_bin[bin[k]] = k;
  for (int i = 0; i < 16; i++)
    for (int tid = 0; tid < FFT_THREADS; tid++)
    {
      const int c = carrier_of_bin[fft_out_index(tid, i >> 3, i & 7)];
      if (c >= 0 && sw[i * FFT_THREADS + tid] != sr[c]) { printf("writer / reader slot mismatch\n"); return 1; }
      if (sw[i * FFT_THREADS + tid] >= 2048) { printf("slot out of range\n"); return 1; }
    }
  for (int i = 0; i < 16; i++)
    for (int h16 = 0; h16 < 8; h16++)
    {
      std::set<int> banks;
      for (int l = 0; l < 16; l++) banks.insert(sw[i * FFT_THREADS + 16 * h16 + l] % 16);
      if (banks.size() != 16) { printf("write conflict i=%d half=%d: %zu banks\n", i, h16, banks.size()); return 1; }
    }
  for (int it = 0; it < 6; it++)
    for (int b = 0; b < 2; b++)
      for (int h16 = 0; h16 < 8; h16++)
      {
        std::set<int> banks;
        for (int l = 0; l < 16; l++) banks.insert(sr[2 * (it * FFT_THREADS + 16 * h16 + l) + b] % 16);
        if (banks.size() != 16) { printf("read conflict it=%d b=%d half=%d: %zu banks\n", it, b, h16, banks.size()); return 1; }
      }
  printf("layout ok\n");
  return 0;
}
