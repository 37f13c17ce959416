/* dab_oracle — CPU restatement of the reference's DAB Mode-I hot path (TEST INFRASTRUCTURE).
 *
 * Plain C, single threaded, written from the reference's behaviour (each function cites the
 * file:line under /root/reference/src it follows). It exists so the CUDA path can be checked on a
 * box where /root/reference does not exist. It is pinned against the reference itself
 * (oracle/_ref/libdabref.so, the unmodified reference sources) by tests/test_oracle_vs_ref.py and
 * against committed vectors in tests/golden/ that were generated from oracle/_ref.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 * The product (dabstar_b200/) never does.
 *
 * The entry points mirror oracle/ref_build/dabref.h one to one (prefix dabo_ instead of dabref_)
 * so the same test code can drive either library.
 */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* tables */
void dabo_freq_interleaver(int16_t out[1536]);
void dabo_phase_table(float out_re_im[4096]);
void dabo_fft2048(const float * in, float * out, int sign);
void dabo_prbs(uint8_t * out, int n);                       /* energy dispersal sequence */
int  dabo_fic_addresses(int32_t * addr, int cap);           /* FIC depuncture map, returns 2304 */

/* sample formats (raw_reader.cpp:66-70,155-158; xml_reader.cpp:254-372) */
void dabo_convert_u8(const uint8_t * in, float * out_re_im, int64_t n_samples);
void dabo_convert_i16(const int16_t * in, float * out_re_im, int64_t n_samples);
/* XmlReader::readElements_IQ / _QI / _I / _Q (xml_filereader/xml_reader.cpp:254-800) for every container, byte order and
 * IQ order. container: 0 int8, 1 uint8, 2 int16, 3 int24, 4 int32, 5 float32; order: 0 IQ, 1 QI, 2 I_Only, 3 Q_Only;
 * bits = bitsperChannel (scaler 2^(bits-1)). The reader's wrong-byte bugs (:309, :421, :690) are not restated. Returns 0, or
 * -1 for an unknown format. */
int dabo_convert_samples(const uint8_t * in, int container, int bits, int msb_first, int order, int64_t n_samples, float * out_re_im);
/* (containers 6 / 7: 8-bit and 32-bit PCM of a RIFF/WAVE file as libsndfile's float read scales them, IQ order only) */
/* SampleReader's DC / IQ-imbalance correction (sample_reader.cpp:216-243), serial float as the reference's scalar build runs it.
 * state: { meanI, meanQ, meanII, meanQQ, meanIQ } (a fresh reader: 0, 0, 1, 1, 0), in and out. */
void dabo_dc_iq_correct(const float * in_re_im, int64_t n, int do_iq, float state[5], float * out_re_im);
/* Linear-interpolation resampling to 2.048 MS/s of XmlReader (xml_reader.cpp:70-76,212-231) and WavReader
 * (wav_reader.cpp:66-83,196-211); out must hold 2048 samples per 1 ms of input. Returns the samples written. */
int64_t dabo_resample_xml(const float * in_re_im, int64_t n_in, int sample_rate, float * out_re_im);
int64_t dabo_resample_wav(const float * in_re_im, int64_t n_in, int sample_rate, float * out_re_im);

/* channel decoding */
void dabo_viterbi(const int16_t * in, int frame_bits, uint8_t * out);
void dabo_viterbi_ber(const int16_t * in, const uint8_t * punct, const uint8_t * out_bits, int frame_bits, int * bits, int * errors);
int  dabo_protection(int short_form, int bit_rate, int prot_level, const int16_t * in, int in_len, uint8_t * out);
int  dabo_protection_addresses(int short_form, int bit_rate, int prot_level, int32_t * addr, int cap);
int  dabo_check_crc_bits(const uint8_t * bits, int n);
uint16_t dabo_calc_crc(const uint8_t * bytes, int n);

void * dabo_fic_new(void);
void   dabo_fic_free(void * h);
void   dabo_fic_process_block(void * h, const int16_t * soft3072, int sym_idx);
void   dabo_fic_get(void * h, uint8_t bits[3072], uint8_t valid[4], int * ratio_percent, int * ber_bits, int * ber_errors);

void * dabo_backend_new(int sub_ch_id, int start_cu, int size_cu, int short_form, int prot_level, int bit_rate);
void   dabo_backend_free(void * h);
int    dabo_backend_process(void * h, const int16_t * fragment, uint8_t * out);

/* DAB+ outer code (dab_outer.c): ReedSolomon(8, 0435, 0, 1, 10).dec(in, out, 135) (backend/reed_solomon.cpp:140-260), FirecodeChecker
 * (backend/firecode_checker.cpp), check_crc_bytes (backend/crc.cpp:89-96), Mp4Processor's super-frame handling (mp4processor.cpp:95-334) */
int  dabo_rs_decode(const uint8_t * in120, uint8_t * out110);            /* returns corrections, 0, or -1 */
void dabo_rs_encode(const uint8_t * in110, uint8_t * out120);            /* test-vector encoder */
int  dabo_firecode_check(const uint8_t * x11);
int  dabo_firecode_check_and_correct(uint8_t * x11);
void dabo_firecode_syndrome_table(uint16_t * out65536);
int  dabo_check_crc_bytes(const uint8_t * msg, int len);
typedef struct
{
  int32_t first_frame;      /* logical frame (CIF) index of the super-frame's first block */
  int32_t ok;               /* Reed-Solomon + Fire code accepted (Mp4Processor::_process_reed_solomon_frame) */
  int32_t rs_errors, rs_corrections, fc_corrected;
  int32_t dac_rate, sbr_flag, aac_channel_mode, ps_flag, mpeg_surround;
  int32_t num_aus;
  int32_t au_start[7];
  int32_t au_state[6];      /* 0 bad length, 1 CRC good, 2 CRC error */
} dabo_superframe;
int  dabo_dabplus_run(const uint8_t * frame_bits, int bit_rate, int n_frames, dabo_superframe * out, int cap, uint8_t * payload);

/* OFDM */
void * dabo_ofdm_new(int soft_bit_type);
void   dabo_ofdm_free(void * h);
void   dabo_ofdm_reset(void * h);
void   dabo_ofdm_store_reference_symbol_0(void * h, const float * fft);
void   dabo_ofdm_store_null_symbol_without_tii(void * h, const float * fft);
void   dabo_ofdm_decode_symbol(void * h, const float * fft, int sym_idx, float phase_corr, float clock_err, int16_t * out3072);
void   dabo_ofdm_get_state(void * h, int which, float * out);
/* SLcdData figures from the current state: { MER dB, SNR dB, mMeanValue, mMeanPowerOvrAll, noise power, sqrt(mMeanSigmaSqFreqCorr) } */
void   dabo_ofdm_quality(void * h, float out[6]);
/* last SLcdData emitted at the reference's cadence { CurOfdmSymbolNo, MeanSigmaSqFreqCorr, SNR, MER, TestData1, TestData2 }; returns the emission count */
int    dabo_ofdm_lcd(void * h, float out[6]);

void * dabo_phaseref_new(void);
void   dabo_phaseref_free(void * h);
int    dabo_phaseref_correlate(void * h, const float * samples2048, float threshold, int strongest_peak);
int    dabo_phaseref_estimate_offset(void * h, const float * fft2048);

/* whole chain (DabProcessor::run, dab_processor.cpp:110-189) */
typedef struct
{
  int   soft_bit_type;
  float threshold;
  int   strongest_peak;
  int   scan_mode;
  int   tap_soft_bits;
  int   tap_fft;
  int   n_subch;
  const int32_t * subch;     /* n_subch x 7: subChId,startCU,sizeCU,shortForm,protLevel,bitRate,startFrame */
  const char * eti_path;     /* NULL = ETI generator off; else the ETI-NI stream is also written to this file (as dabref does) */
  int   eti_cif_hi, eti_cif_lo; /* IFibDecoder::get_cif_count(hi, lo) as sampled at symbol 4 of every frame (the dabref stub: 0, 0);
                                   eti_cif_hi < 0: the counter of the last FIG 0/0 received, as a real FIB decoder reports it */
  int   track_cif;           /* 1: get_cif_count() follows the recording's own FIG 0/0 (a real FIB decoder), so the null symbols with
                                (count & 7) >= 4 are TII symbols and do not update the null power (dab_processor.cpp:273-285); 0: the count stays 0 */
} dabo_chain_cfg;

typedef struct
{
  int64_t sym0_pos;
  int32_t start_index;
  float   fbb_sym0;
  float   fbb_data;
  float   fbb_null;
  float   fsync;
  float   phase_cp;
  float   clock_err;
  int32_t fic_ratio_before;
  int32_t fic_ratio_after;
  uint8_t fic_valid[4];
} dabo_frame_info;

void *  dabo_chain_run(const float * iq_re_im, int64_t n_samples, const dabo_chain_cfg * cfg);
void    dabo_chain_free(void * h);
int     dabo_chain_n_frames(void * h);
void    dabo_chain_frame_info(void * h, int frame, dabo_frame_info * out);
void    dabo_chain_fib_bits(void * h, int frame, uint8_t out[3072]);
int     dabo_chain_soft_bits(void * h, int frame, int16_t * out);
int     dabo_chain_fft(void * h, int frame, float * out);
int     dabo_chain_n_good_fibs(void * h);
int64_t dabo_chain_msc_size(void * h, int sub_ch_id);
int64_t dabo_chain_msc_copy(void * h, int sub_ch_id, uint8_t * out, int64_t cap);
/* ETI-NI frames (6144 bytes each) produced by the run (eti_generator.cpp:90-204); only in the restatement, dabref writes the file */
int64_t dabo_chain_eti_size(void * h);
int64_t dabo_chain_eti_copy(void * h, uint8_t * out, int64_t cap);
void    dabo_chain_counters(void * h, int64_t out[8]);
void    dabo_chain_quality(void * h, float out[6]); /* dabo_ofdm_quality of the chain's decoder at the end of the run */
int     dabo_chain_lcd(void * h, float out[6]);     /* dabo_ofdm_lcd of the chain's decoder */
double  dabo_chain_seconds(void * h);

/* TII detector (ofdm/tii_detector.cpp; restatement in dab_tii.c) */
typedef struct { int32_t main_id, sub_id; float strength, phase_deg; int32_t non_etsi; } dabo_tii_result;
void * dabo_tii_new(void);
void   dabo_tii_free(void * h);
void   dabo_tii_reset(void * h);
void   dabo_tii_set_collisions(void * h, int on, int sub_id);
void   dabo_tii_add(void * h, const float * fft2048);                                  /* add_to_tii_buffer */
int    dabo_tii_process(void * h, int threshold_db, dabo_tii_result * out, int cap);  /* process_tii_data; returns the count */
void   dabo_tii_decoded(void * h, float * out_re_im);                                  /* mDecodedBufferArr, 768 complex */
void   dabo_tii_tables(void * h, uint8_t pattern[70], uint8_t phase_corr[768]);        /* the two derived constant tables */

#ifdef __cplusplus
}
#endif
