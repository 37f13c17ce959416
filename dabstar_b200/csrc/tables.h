// tables.h — host builders of the constant tables (see tables.cu for the reference citations).
#pragma once
#include "common.cuh"
#include "viterbi.cuh"

namespace dab
{
void host_freq_interleaver(int16_t * bin_signed);  // 1536 entries in [-768, 768] \ {0}
void host_phase_table(float2 * prs);               // 2048 entries, fft order
void host_prbs(uint8_t * out, int n);
void host_w2048(float2 * w);
// Conflict-free staging layout of k_fft_frames' epilogue (tables.cu): slot (in float2 units) written by output i of thread tid,
// and slot of every nominal carrier
void host_fft_epilogue_layout(const int16_t * bin_of_k, uint16_t * slot_w /* [16][128] */, uint16_t * slot_r /* [1536] */);
VitProfile make_fic_profile();
VitProfile make_identity_profile(int n_bits);
bool make_msc_profile(int short_form, int bit_rate, int prot_level, VitProfile & p);
int profile_addresses(const VitProfile & p, int32_t * addr, int cap);
} // namespace dab
