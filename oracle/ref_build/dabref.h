/* C interface of oracle/_ref/libdabref.so: the UNMODIFIED reference hot-path sources
 * (/root/reference/src/...) compiled against stub headers, driven through these entry points.
 * Test infrastructure: used by tests/, tools/make_golden.py and bench.py's CPU-baseline legs only.
 * Not reentrant (the reference keeps file-scope state, viterbi_spiral.cpp:40-41): one call at a time
 * per process.
 */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* tables */
void dabref_freq_interleaver(int16_t out[1536]);                 /* FreqInterleaver::map_k_to_fft_bin */
void dabref_phase_table(float out_re_im[4096]);                  /* PhaseTable::mRefTable */
void dabref_fft2048(const float * in, float * out, int sign);    /* the FFTW shim itself */

/* channel decoding */
/* DAB+ outer code: the reference's ReedSolomon(8, 0435, 0, 1, 10), FirecodeChecker and check_crc_bytes objects */
int  dabref_rs_decode(const uint8_t * in120, uint8_t * out110);
void dabref_rs_encode(const uint8_t * in110, uint8_t * out120);
int  dabref_firecode_check(const uint8_t * x11);
int  dabref_firecode_check_and_correct(uint8_t * x11);
int  dabref_check_crc_bytes(const uint8_t * msg, int len);
void dabref_viterbi(const int16_t * in, int frame_bits, uint8_t * out);                 /* ViterbiSpiral::deconvolve */
void dabref_viterbi_ber(const int16_t * in, const uint8_t * punct, const uint8_t * out_bits, int frame_bits, int * bits, int * errors);
int  dabref_protection(int short_form, int bit_rate, int prot_level, const int16_t * in, int in_len, uint8_t * out);
int  dabref_protection_addresses(int short_form, int bit_rate, int prot_level, int32_t * addr, int cap); /* returns count */
int  dabref_check_crc_bits(const uint8_t * bits, int n);
uint16_t dabref_calc_crc(const uint8_t * bytes, int n);

void * dabref_fic_new(void);
void   dabref_fic_free(void * h);
void   dabref_fic_process_block(void * h, const int16_t * soft3072, int sym_idx);
void   dabref_fic_get(void * h, uint8_t bits[3072], uint8_t valid[4], int * ratio_percent, int * ber_bits, int * ber_errors);

void * dabref_backend_new(int sub_ch_id, int start_cu, int size_cu, int short_form, int prot_level, int bit_rate);
void   dabref_backend_free(void * h);
int    dabref_backend_process(void * h, const int16_t * fragment, uint8_t * out); /* 1 if a logical frame was emitted */

/* OFDM */
void * dabref_ofdm_new(int soft_bit_type);
void   dabref_ofdm_free(void * h);
void   dabref_ofdm_reset(void * h);
void   dabref_ofdm_store_reference_symbol_0(void * h, const float * fft);
void   dabref_ofdm_store_null_symbol_without_tii(void * h, const float * fft);
void   dabref_ofdm_decode_symbol(void * h, const float * fft, int sym_idx, float phase_corr, float clock_err, int16_t * out3072);
/* which: 0 integAbsPhase[1536] 1 stdDevSq[1536] 2 meanPower[1536] 3 meanSigmaSq[1536] 4 nullPower[2048] 5 {meanValue, meanPowerOvrAll} */
void   dabref_ofdm_get_state(void * h, int which, float * out);

void * dabref_phaseref_new(void);
void   dabref_phaseref_free(void * h);
int    dabref_phaseref_correlate(void * h, const float * samples2048, float threshold, int strongest_peak);
int    dabref_phaseref_estimate_offset(void * h, const float * fft2048);

/* whole chain: DabProcessor::run over an in-memory sample source */
typedef struct
{
  int   soft_bit_type;       /* 0 = SOFTDEC1 (default) */
  float threshold;           /* 3.0 */
  int   strongest_peak;      /* 0 */
  int   scan_mode;           /* 1: FIC only (dab_processor.cpp:357) */
  int   tap_soft_bits;       /* keep 75x3072 i16 per frame */
  int   tap_fft;             /* keep 77x2048 cf32 per frame */
  int   n_subch;
  const int32_t * subch;     /* n_subch x 7: subChId,startCU,sizeCU,shortForm,protLevel,bitRate,startFrame */
  const char * eti_path;     /* NULL = ETI generator off */
  int   eti_cif_hi, eti_cif_lo; /* (layout shared with the restatement's configuration; not used here: get_cif_count(hi, lo) gives 0, 0) */
  int   track_cif;           /* 1: the stub FIB decoder's get_cif_count() follows the FIG 0/0 of the FIBs FicDecoder hands it; 0: always 0 */
} dabref_chain_cfg;

typedef struct
{
  int64_t sym0_pos;          /* stream index of the first sample of symbol 0's useful part */
  int32_t start_index;       /* PhaseReference peak index */
  float   fbb_sym0;          /* mFreqOffsBBHz while symbol 0 was read */
  float   fbb_data;          /* ... while symbols 1..75 were read (after coarse AFC) */
  float   fbb_null;          /* ... while the null symbol was read (after fine AFC) */
  float   fsync;             /* mFreqOffsSyncSymb after the fine update */
  float   phase_cp;          /* clamped arg of the CP correlation */
  float   clock_err;         /* mClockErrHz used by decode_symbol in this frame */
  int32_t fic_ratio_before;  /* get_fic_decode_ratio_percent() at the start of the frame */
  int32_t fic_ratio_after;
  uint8_t fic_valid[4];
} dabref_frame_info;

/* TiiDetector (ofdm/tii_detector.h:30-45): null-symbol spectra in, transmitter identifications out */
typedef struct { int32_t main_id, sub_id; float strength, phase_deg; int32_t non_etsi; } dabref_tii_result;
void * dabref_tii_new(void);
void   dabref_tii_free(void * h);
void   dabref_tii_reset(void * h);
void   dabref_tii_set_collisions(void * h, int on, int sub_id);
void   dabref_tii_add(void * h, const float * fft2048);                                      /* add_to_tii_buffer */
int    dabref_tii_process(void * h, int threshold_db, dabref_tii_result * out, int cap);    /* process_tii_data; returns the count */

/* SampleReader (ofdm/sample_reader.h:60-70) with DC (+ IQ) correction switched on, 0 Hz mixing, over n samples */
void    dabref_dc_iq_correct(const float * in_re_im, int64_t n, int do_iq, float * out_re_im, float dc_offset[2]);
void *  dabref_chain_run(const float * iq_re_im, int64_t n_samples, const dabref_chain_cfg * cfg);
void    dabref_chain_free(void * h);
int     dabref_chain_n_frames(void * h);
void    dabref_chain_frame_info(void * h, int frame, dabref_frame_info * out);
void    dabref_chain_fib_bits(void * h, int frame, uint8_t out[3072]);
int     dabref_chain_soft_bits(void * h, int frame, int16_t * out /* 75*3072 */);
int     dabref_chain_fft(void * h, int frame, float * out /* 77*2048*2 */);
int     dabref_chain_n_good_fibs(void * h);
int64_t dabref_chain_msc_size(void * h, int sub_ch_id);
int64_t dabref_chain_msc_copy(void * h, int sub_ch_id, uint8_t * out, int64_t cap);
void    dabref_chain_counters(void * h, int64_t out[8]); /* dip found, no dip, single-sample reads, samples consumed, last snr*1000, last mer*1000, 0, 0 */
double  dabref_chain_seconds(void * h);                   /* steady_clock around DabProcessor::run() */

#ifdef __cplusplus
}
#endif
