// oracle/_ref harness (TEST INFRASTRUCTURE): drives the unmodified reference classes through dabref.h.
// Nothing here is product code; the CUDA path never links or loads this.
#include "harness_internal.h"
#include "dabref.h"
#include "reed_solomon.h"
#include "firecode_checker.h"
#include "crc.h"

#include <chrono>
#include <cstring>
#include <functional>
#include <map>

Hooks * gHooks = nullptr;

namespace
{
DabRadio gRadio;

struct ScopedHooks
{
  explicit ScopedHooks(Hooks * h) : mPrev(gHooks) { gHooks = h; }
  ~ScopedHooks() { gHooks = mPrev; }
  Hooks * mPrev;
};

const TArrayTu & as_tu(const float * p) { return *reinterpret_cast<const TArrayTu *>(p); }
}

// ------------------------------------------------------------------------------------------------ tables
extern "C" void dabref_freq_interleaver(int16_t out[1536])
{
  FreqInterleaver fi;
  for (int k = 0; k < cK; k++) out[k] = fi.map_k_to_fft_bin((i16)k);
}

extern "C" void dabref_phase_table(float out[4096])
{
  PhaseTable pt;
  memcpy(out, pt.mRefTable.data(), sizeof(cf32) * cTu);
}

extern "C" void dabref_fft2048(const float * in, float * out, int sign)
{
  static TArrayTu bi, bo;
  fftwf_plan p = fftwf_plan_dft_1d(cTu, (fftwf_complex *)bi.data(), (fftwf_complex *)bo.data(), sign, FFTW_ESTIMATE);
  memcpy(bi.data(), in, sizeof(cf32) * cTu);
  fftwf_execute(p);
  memcpy(out, bo.data(), sizeof(cf32) * cTu);
  fftwf_destroy_plan(p);
}

// ------------------------------------------------------------------------------------------------ DAB+ outer code
// The reference's own ReedSolomon(8, 0435, 0, 1, 10) and FirecodeChecker as Mp4Processor uses them (mp4processor.cpp:52,184-241).
extern "C" int dabref_rs_decode(const uint8_t * in120, uint8_t * out110)
{
  static ReedSolomon rs(8, 0435, 0, 1, 10);
  return rs.dec(in120, out110, 135);
}
extern "C" void dabref_rs_encode(const uint8_t * in110, uint8_t * out120)
{
  static ReedSolomon rs(8, 0435, 0, 1, 10);
  rs.enc(in110, out120, 135);
}
extern "C" int dabref_firecode_check(const uint8_t * x11)
{
  static FirecodeChecker fc;
  return fc.check(x11) ? 1 : 0;
}
extern "C" int dabref_firecode_check_and_correct(uint8_t * x11)
{
  static FirecodeChecker fc;
  return fc.check_and_correct_6bits(x11) ? 1 : 0;
}
extern "C" int dabref_check_crc_bytes(const uint8_t * msg, int len) { return check_crc_bytes(msg, len) ? 1 : 0; }

// The reference's own file-reader threads: XmlReader (sample format conversion of every container / byte order / channel order,
// 1 ms linear-interpolation resampling) and WavReader (resampling of libsndfile's float frames). The threads pace themselves in
// real time (1 ms of sleep per millisecond of recording): keep the inputs to a fraction of a second.
static u32 pow2_at_least(long long n) { u32 p = 1024; while ((long long)p < n) p <<= 1; return p; }
extern "C" int64_t dabref_xml_reader_run(const uint8_t * bytes, int64_t n_bytes, int sample_rate, int bits, const char * container, const char * byte_order,
                                         const char * iq_order, int64_t n_samples, float * out, int64_t cap)
{
  FILE * f = fmemopen(const_cast<uint8_t *>(bytes), (size_t)n_bytes, "rb");
  if (!f) return -1;
  bool ok = false;
  XmlDescriptor fd(nullptr, &ok);
  fd.sampleRate = sample_rate;
  fd.nrChannels = 2;
  fd.bitsperChannel = bits;
  fd.container = container;
  fd.byteOrder = byte_order;
  fd.iqOrder = iq_order;
  XmlFileReader parent;
  parent.samplesToRead = n_samples;
  parent.mFileLength = n_bytes;
  RingBuffer<cf32> rb(pow2_at_least((n_samples / std::max(1, sample_rate / 1000) + 4) * 2048));
  {
    XmlReader reader(&parent, f, &fd, 0, &rb); // the constructor starts the thread; the QThread stand-in runs it to its end here
  }
  fclose(f);
  const int64_t n = std::min<int64_t>(rb.get_ring_buffer_read_available(), cap);
  rb.get_data_from_ring_buffer(reinterpret_cast<cf32 *>(out), (i32)n);
  return n;
}
extern "C" int64_t dabref_wav_reader_run(const float * frames, int64_t n_frames, int sample_rate, float * out, int64_t cap)
{
  SNDFILE snd{ frames, (sf_count_t)n_frames, 0 };
  WavFileHandler parent;
  RingBuffer<cf32> rb(pow2_at_least((n_frames / std::max(1, sample_rate / 1000) + 4) * 2048 + 4 * 32768));
  {
    WavReader reader(&parent, &snd, &rb, sample_rate);
    reader.start_reader();
  }
  const int64_t n = std::min<int64_t>(rb.get_ring_buffer_read_available(), cap);
  rb.get_data_from_ring_buffer(reinterpret_cast<cf32 *>(out), (i32)n);
  return n;
}

// The reference's own FibDecoder (decoder/fib_decoder*.cpp, fib_config_fig0/1.cpp) fed with CRC-good FIBs: the multiplex
// configuration it extracts (FIG 0/0 CIF counter and EId, FIG 0/1 sub-channel organisation, FIG 0/2 service components).
extern "C" void * dabref_fibdec_new(void)
{
  auto * d = new FibDecoder(nullptr);
  d->connect_channel();
  return d;
}
extern "C" void dabref_fibdec_free(void * h) { delete static_cast<FibDecoder *>(h); }
// bits: 256 bytes, one bit each (30 bytes of FIGs + CRC), as FicDecoder hands them over (fic_decoder.cpp:239-247)
extern "C" void dabref_fibdec_process(void * h, const uint8_t * bits, int fic_no)
{
  std::array<std::byte, cFibSizeVitOut> a;
  for (int i = 0; i < cFibSizeVitOut; i++) a[i] = (std::byte)bits[i];
  static_cast<FibDecoder *>(h)->process_FIB(a, (u16)fic_no);
}
extern "C" void dabref_fibdec_cif_count(void * h, int * hi, int * lo)
{
  i16 a = 0, b = 0;
  static_cast<FibDecoder *>(h)->get_cif_count(&a, &b);
  *hi = a;
  *lo = b;
}
extern "C" int dabref_fibdec_eid(void * h) { return static_cast<FibDecoder *>(h)->get_EId(); }
// out: in_use, id, start_cu, short form (uepFlag), protection level, size in CU, bit rate
extern "C" void dabref_fibdec_subch(void * h, int sub_ch_id, int out[7])
{
  SChannelData d{};
  static_cast<FibDecoder *>(h)->get_sub_channel_info(&d, sub_ch_id);
  out[0] = d.in_use ? 1 : 0; out[1] = d.id; out[2] = d.start_cu; out[3] = d.uepFlag; out[4] = d.protlev; out[5] = d.size; out[6] = d.bitrate;
}
extern "C" int dabref_fibdec_subch_list(void * h, int8_t * out, int cap)
{
  const std::vector<i8> v = static_cast<FibDecoder *>(h)->get_sub_channel_id_list();
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
// Service components as FIG 0/2 filed them in the CURRENT configuration: per component SId, TMId, sub-channel id (or SCId for
// packet mode), ASCTy/DSCTy, primary flag. Returns the number of components (out receives up to cap rows of 6 ints).
extern "C" int dabref_fibdec_components(void * h, int * out, int cap)
{
  auto * d = static_cast<FibDecoder *>(h);
  int n = 0;
  for (const auto & e : d->mpFibConfigFig0Curr->Fig0s2_BasicService_ServiceCompDefVec)
  {
    if (n < cap)
    {
      int * o = out + 6 * n;
      const auto & c = e.ServiceComp_C;
      o[0] = (int)e.get_SId(); o[1] = c.TMId; o[2] = c.TMId == 3 ? (int)c.TMId11.SCId : (int)c.TMId00.SubChId;
      o[3] = c.TMId == 3 ? 0 : (int)c.TMId00.ASCTy; o[4] = c.PS_Flag; o[5] = e.ServiceComp_C_index;
    }
    n++;
  }
  return n;
}

// The reference's own Mp4Processor (backend/audio/mp4processor.cpp: super-frame synchronisation, RS repair, Fire code, AU borders
// and CRCs) fed with logical frames; its AAC decoder is a stand-in that records the calls (stubs/faad_decoder.h).
std::vector<DabrefAacEvent> * gDabrefAacSink = nullptr;
struct Mp4Box
{
  std::vector<DabrefAacEvent> events;
  RingBuffer<i16> audio{ 64 * 32768 };
  RingBuffer<u8> frames{ 2 * 32768 };
  std::unique_ptr<Mp4Processor> proc;
  int bitRate = 0;
};
extern "C" void * dabref_mp4_new(int bit_rate)
{
  auto * b = new Mp4Box;
  b->bitRate = bit_rate;
  b->proc.reset(new Mp4Processor(nullptr, (i16)bit_rate, &b->audio, &b->frames));
  return b;
}
extern "C" void dabref_mp4_free(void * h) { delete static_cast<Mp4Box *>(h); }
// frame_bits: n_frames x 24 * bit_rate bytes (one bit per byte). Returns the number of decoder events so far.
extern "C" int dabref_mp4_add_frames(void * h, const uint8_t * frame_bits, int n_frames)
{
  auto * b = static_cast<Mp4Box *>(h);
  gDabrefAacSink = &b->events;
  const size_t len = 24 * (size_t)b->bitRate;
  for (int i = 0; i < n_frames; i++)
  {
    std::vector<u8> v(frame_bits + (size_t)i * len, frame_bits + (size_t)(i + 1) * len);
    b->proc->add_to_frame(v);
    // the "AAC dump" ring buffer is never drained here: keep it from filling up
    b->frames.flush_ring_buffer();
  }
  gDabrefAacSink = nullptr;
  return (int)b->events.size();
}
// Event i: kind (1 access unit / 0 concealment), value (stream parameters / samples), bytes written to `out` (up to cap).
extern "C" int dabref_mp4_event(void * h, int i, int * kind, int * value, uint8_t * out, int cap)
{
  auto * b = static_cast<Mp4Box *>(h);
  if (i < 0 || i >= (int)b->events.size()) return -1;
  const DabrefAacEvent & e = b->events[(size_t)i];
  *kind = e.kind;
  *value = e.value;
  const int n = (int)std::min<size_t>(e.bytes.size(), (size_t)std::max(cap, 0));
  if (n > 0) memcpy(out, e.bytes.data(), (size_t)n);
  return (int)e.bytes.size();
}

// ------------------------------------------------------------------------------------------------ channel decoding
extern "C" void dabref_viterbi(const int16_t * in, int frame_bits, uint8_t * out)
{
  ViterbiSpiral v((i16)frame_bits, true);
  v.deconvolve(in, out);
}

extern "C" void dabref_viterbi_ber(const int16_t * in, const uint8_t * punct, const uint8_t * out_bits, int frame_bits, int * bits, int * errors)
{
  ViterbiSpiral v((i16)frame_bits, true);
  i32 b = 0, e = 0;
  v.calculate_BER(in, const_cast<u8 *>(punct), out_bits, b, e);
  *bits = b;
  *errors = e;
}

static Protection * make_protection(int short_form, int bit_rate, int prot_level)
{
  if (short_form) return new UepProtection((i16)bit_rate, (i16)prot_level);
  return new EepProtection((i16)bit_rate, (i16)prot_level);
}

extern "C" int dabref_protection(int short_form, int bit_rate, int prot_level, const int16_t * in, int in_len, uint8_t * out)
{
  std::unique_ptr<Protection> p(make_protection(short_form, bit_rate, prot_level));
  return p->deconvolve(in, in_len, out) ? 1 : 0;
}

extern "C" int dabref_protection_addresses(int short_form, int bit_rate, int prot_level, int32_t * addr, int cap)
{
  std::unique_ptr<Protection> p(make_protection(short_form, bit_rate, prot_level));
  int n = 0;
  for (i16 * a : p->viterbiBlockAddresses)
  {
    if (n < cap) addr[n] = (int32_t)(a - p->viterbiBlock.data());
    n++;
  }
  return n;
}

extern "C" int dabref_check_crc_bits(const uint8_t * bits, int n) { return check_CRC_bits(bits, n) ? 1 : 0; }
extern "C" uint16_t dabref_calc_crc(const uint8_t * bytes, int n) { return calc_crc(bytes, n); }

struct FicBox
{
  Hooks hooks;
  FicDecoder fic{ &gRadio };
};

extern "C" void * dabref_fic_new(void)
{
  auto * b = new FicBox;
  b->fic.restart();
  return b;
}
extern "C" void dabref_fic_free(void * h) { delete static_cast<FicBox *>(h); }
extern "C" void dabref_fic_process_block(void * h, const int16_t * soft, int sym_idx)
{
  auto * b = static_cast<FicBox *>(h);
  ScopedHooks s(&b->hooks);
  std::vector<i16> v(soft, soft + c2K);
  b->fic.process_block(v, sym_idx);
}
extern "C" void dabref_fic_get(void * h, uint8_t bits[3072], uint8_t valid[4], int * ratio_percent, int * ber_bits, int * ber_errors)
{
  auto * b = static_cast<FicBox *>(h);
  bool v[4];
  b->fic.get_fib_bits(bits, v);
  for (int i = 0; i < 4; i++) valid[i] = v[i] ? 1 : 0;
  *ratio_percent = b->fic.get_fic_decode_ratio_percent();
  *ber_bits = b->fic.mFicBits;
  *ber_errors = b->fic.mFicErrors;
}

struct BackendBox : Hooks
{
  SAudioData desc;
  std::unique_ptr<Backend> backend;
  std::vector<u8> last;
  bool emitted = false;
  void on_msc_frame(int, const u8 * d, int n) override { last.assign(d, d + n); emitted = true; }
};

extern "C" void * dabref_backend_new(int sub_ch_id, int start_cu, int size_cu, int short_form, int prot_level, int bit_rate)
{
  auto * b = new BackendBox;
  b->desc.isDefined = true;
  b->desc.SId = 0x1000 + sub_ch_id;
  b->desc.SubChId = (i16)sub_ch_id;
  b->desc.CuStartAddr = (i16)start_cu;
  b->desc.CuSize = (i16)size_cu;
  b->desc.shortForm = short_form != 0;
  b->desc.protLevel = (i16)prot_level;
  b->desc.bitRate = (i16)bit_rate;
  b->desc.ASCTy = 077;
  ScopedHooks s(b);
  b->backend = std::make_unique<Backend>(&gRadio, &b->desc, nullptr, nullptr, nullptr, EProcessFlag::Primary);
  return b;
}
extern "C" void dabref_backend_free(void * h) { delete static_cast<BackendBox *>(h); }
extern "C" int dabref_backend_process(void * h, const int16_t * fragment, uint8_t * out)
{
  auto * b = static_cast<BackendBox *>(h);
  ScopedHooks s(b);
  b->emitted = false;
  b->backend->process(fragment, b->desc.CuSize * 64);
  if (b->emitted) memcpy(out, b->last.data(), b->last.size());
  return b->emitted ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ OFDM
static RingBuffer<cf32> gIqRing{ 2 * 1536 };
static RingBuffer<f32> gCarrRing{ 2 * 1536 };

extern "C" void * dabref_ofdm_new(int soft_bit_type)
{
  auto * d = new OfdmDecoder(&gRadio, &gIqRing, &gCarrRing);
  d->set_soft_bit_gen_type((ESoftBitType)soft_bit_type);
  return d;
}
extern "C" void dabref_ofdm_free(void * h) { delete static_cast<OfdmDecoder *>(h); }
extern "C" void dabref_ofdm_reset(void * h) { static_cast<OfdmDecoder *>(h)->reset(); }
extern "C" void dabref_ofdm_store_reference_symbol_0(void * h, const float * fft) { static_cast<OfdmDecoder *>(h)->store_reference_symbol_0(as_tu(fft)); }
extern "C" void dabref_ofdm_store_null_symbol_without_tii(void * h, const float * fft) { static_cast<OfdmDecoder *>(h)->store_null_symbol_without_tii(as_tu(fft)); }
extern "C" void dabref_ofdm_decode_symbol(void * h, const float * fft, int sym_idx, float phase_corr, float clock_err, int16_t * out)
{
  std::vector<i16> bits(c2K);
  static_cast<OfdmDecoder *>(h)->decode_symbol(as_tu(fft), (u16)sym_idx, phase_corr, clock_err, bits);
  memcpy(out, bits.data(), sizeof(i16) * c2K);
}
extern "C" void dabref_ofdm_get_state(void * h, int which, float * out)
{
  auto * d = static_cast<OfdmDecoder *>(h);
  switch (which)
  {
  case 0: memcpy(out, d->mIntegAbsPhaseVector.data(), sizeof(f32) * cK); break;
  case 1: memcpy(out, d->mStdDevSqPhaseVector.data(), sizeof(f32) * cK); break;
  case 2: memcpy(out, d->mMeanPowerVector.data(), sizeof(f32) * cK); break;
  case 3: memcpy(out, d->mMeanSigmaSqVector.data(), sizeof(f32) * cK); break;
  case 4: memcpy(out, d->mMeanNullPowerWithoutTII.data(), sizeof(f32) * cTu); break;
  case 5: out[0] = d->mMeanValue; out[1] = d->mMeanPowerOvrAll; break;
  default: break;
  }
}

struct PhaseRefBox
{
  ProcessParams params;
  PhaseReference ref{ &gRadio, &params };
};
extern "C" void * dabref_phaseref_new(void) { return new PhaseRefBox; }
extern "C" void dabref_phaseref_free(void * h) { delete static_cast<PhaseRefBox *>(h); }
extern "C" int dabref_phaseref_correlate(void * h, const float * samples2048, float threshold, int strongest_peak)
{
  auto * b = static_cast<PhaseRefBox *>(h);
  static TArrayTn buf;
  memcpy(buf.data(), samples2048, sizeof(cf32) * cTu);
  b->ref.set_sync_on_strongest_peak(strongest_peak != 0);
  return b->ref.correlate_with_phase_ref_and_find_max_peak(buf, threshold);
}
extern "C" int dabref_phaseref_estimate_offset(void * h, const float * fft2048)
{
  return static_cast<PhaseRefBox *>(h)->ref.estimate_carrier_offset_from_sync_symbol_0(as_tu(fft2048));
}

// ------------------------------------------------------------------------------------------------ TII detector
extern "C" void * dabref_tii_new(void) { return new TiiDetector(); }
extern "C" void dabref_tii_free(void * h) { delete static_cast<TiiDetector *>(h); }
extern "C" void dabref_tii_reset(void * h) { static_cast<TiiDetector *>(h)->reset(); }
extern "C" void dabref_tii_set_collisions(void * h, int on, int sub_id)
{
  static_cast<TiiDetector *>(h)->set_detect_collisions(on != 0);
  static_cast<TiiDetector *>(h)->set_subid_for_collision_search((u8)sub_id);
}
extern "C" void dabref_tii_add(void * h, const float * fft2048) { static_cast<TiiDetector *>(h)->add_to_tii_buffer(as_tu(fft2048)); }
extern "C" int dabref_tii_process(void * h, int threshold_db, dabref_tii_result * out, int cap)
{
  const std::vector<STiiResult> r = static_cast<TiiDetector *>(h)->process_tii_data((i16)threshold_db);
  for (int i = 0; i < (int)r.size() && i < cap; i++)
    out[i] = dabref_tii_result{ r[i].mainId, r[i].subId, r[i].strength, r[i].phaseDeg, r[i].isNonEtsiPhase ? 1 : 0 };
  return (int)r.size();
}

// ------------------------------------------------------------------------------------------------ whole chain
namespace
{
// In-memory, non-paced sample source (IDeviceHandler, device_handler_if.h:38-62). End of data = throw 20,
// which DabProcessor::run catches as its normal stop condition (dab_processor.cpp:184-188).
class MemoryDevice : public IDeviceHandler
{
public:
  MemoryDevice(const cf32 * ipData, int64_t iN) : mpData(ipData), mN(iN) {}
  bool restartReader(i32) override { return true; }
  void stopReader() override {}
  void setVFOFrequency(i32) override {}
  i32 getVFOFrequency() override { return 0; }
  i32 getSamples(cf32 * oV, i32 iN) override
  {
    if (mFirst) { mFirst = false; if (mOnFirstRead) mOnFirstRead(); }
    if (mPos + iN > mN) throw 20;
    memcpy(oV, mpData + mPos, sizeof(cf32) * iN);
    mPos += iN;
    if (iN == 1) mSingleReads++;
    mPrevRead = mLastRead;
    mLastRead = iN;
    return iN;
  }
  i32 Samples() override { return 1 << 30; }
  void resetBuffer() override {}
  void hide() override {}
  void show() override {}
  bool isHidden() override { return true; }
  QString deviceName() override { return "memory"; }
  bool isFileInput() override { return true; }

  const cf32 * const mpData;
  const int64_t mN;
  int64_t mPos = 0;
  int64_t mSingleReads = 0;
  i32 mLastRead = -1;
  i32 mPrevRead = -1;
  bool mFirst = true;
  std::function<void()> mOnFirstRead;
};

} // namespace

// SampleReader with set_dc_and_iq_correction(true, do_iq) over a whole stream, no mixing (0 Hz): the reader's own
// per-sample loop (sample_reader.cpp:212-281) on an in-memory device.
extern "C" void dabref_dc_iq_correct(const float * in_re_im, int64_t n, int do_iq, float * out_re_im, float dc_offset[2])
{
  MemoryDevice dev(reinterpret_cast<const cf32 *>(in_re_im), n);
  auto rdp = std::make_unique<SampleReader>(nullptr, &dev, nullptr); // (the reader holds a 16 MB oscillator table)
  SampleReader & rd = *rdp;
  rd.set_dc_and_iq_correction(true, do_iq != 0);
  auto buf = std::make_unique<TArrayTn>();
  int64_t pos = 0;
  while (pos < n)
  {
    const i32 cnt = (i32)std::min<int64_t>(2048, n - pos);
    rd.get_samples(*buf, 0, cnt, 0.0f, false);
    memcpy(out_re_im + 2 * pos, buf->data(), sizeof(cf32) * (size_t)cnt);
    pos += cnt;
  }
  const cf32 dc = rd.get_dc_offset();
  dc_offset[0] = real(dc);
  dc_offset[1] = imag(dc);
}

namespace
{
struct FrameRec
{
  dabref_frame_info info{};
  std::vector<u8> fibBits;
  std::vector<i16> soft;
  std::vector<float> fft;
};

struct ChainRun : Hooks
{
  dabref_chain_cfg cfg{};
  std::unique_ptr<MemoryDevice> dev;
  ProcessParams params;
  RingBuffer<cf32> iqRing{ 2 * 1536 };   // scope outputs of OfdmDecoder::decode_symbol (ofdm_decoder.cpp:320-321) need a sink
  RingBuffer<f32> carrRing{ 2 * 1536 };
  std::unique_ptr<DabProcessor> proc;
  std::vector<FrameRec> frames;
  std::map<int, std::vector<u8>> msc;
  int goodFibs = 0;
  int fftInFrame = 0;      // 0 = next mFftPlan execution is symbol 0
  int framesStarted = 0;
  float lastSnr = 0, lastMer = 0;
  int cifCount = 0;
  double seconds = 0;

  void start_backends_for_frame(int frame)
  {
    for (const auto & s : subch)
    {
      if (s.startFrame != frame) continue;
      SAudioData d;
      d.isDefined = true;
      d.SId = 0x1000 + s.subChId;
      d.SubChId = (i16)s.subChId;
      d.CuStartAddr = (i16)s.startCU;
      d.CuSize = (i16)s.sizeCU;
      d.shortForm = s.shortForm != 0;
      d.protLevel = (i16)s.protLevel;
      d.bitRate = (i16)s.bitRate;
      d.ASCTy = 077;
      proc->set_audio_channel(d, nullptr, EProcessFlag::Primary);
    }
  }

  void capture_soft(int symIdx)
  {
    if (!cfg.tap_soft_bits || frames.empty()) return;
    auto & f = frames.back();
    if (f.soft.empty()) f.soft.resize((size_t)75 * c2K);
    memcpy(&f.soft[(size_t)(symIdx - 1) * c2K], proc->mBits.data(), sizeof(i16) * c2K);
  }

  void before_fft(fftwf_plan p) override
  {
    if (!proc || p != proc->mFftPlan) return;
    if (fftInFrame == 0)
    {
      frames.emplace_back();
      auto & f = frames.back();
      f.info.sym0_pos = dev->mPos - cTu;
      f.info.start_index = dev->mLastRead; // second read of _state_eval_sync_symbol has exactly startIndex samples (dab_processor.cpp:409)
      f.info.fbb_sym0 = proc->mFreqOffsBBHz;
      f.info.clock_err = proc->mClockErrHz;
      f.info.fic_ratio_before = proc->mFicHandler.get_fic_decode_ratio_percent();
    }
    else if (fftInFrame == 1)
    {
      frames.back().info.fbb_data = proc->mFreqOffsBBHz;
      frames.back().info.clock_err = proc->mClockErrHz; // coarse AFC may have zeroed it (dab_processor.cpp:218-221)
    }
    if (fftInFrame >= 2) capture_soft(fftInFrame - 1);
  }

  void after_fft(fftwf_plan p, const float * out) override
  {
    if (!proc || p != proc->mFftPlan) return;
    if (cfg.tap_fft && !frames.empty())
    {
      auto & f = frames.back();
      if (f.fft.empty()) f.fft.resize((size_t)77 * cTu * 2);
      memcpy(&f.fft[(size_t)fftInFrame * cTu * 2], out, sizeof(cf32) * cTu);
    }
    fftInFrame = (fftInFrame + 1) % 77;
  }

  // Once per frame, after symbols 1..75 and the fine AFC update, before the null symbol FFT (dab_processor.cpp:274).
  int on_get_cif_count() override
  {
    auto & f = frames.back();
    f.info.fbb_null = proc->mFreqOffsBBHz;
    f.info.fsync = proc->mFreqOffsSyncSymb;
    f.info.phase_cp = proc->mPhaseOffsetCyclPrefRad;
    f.info.fic_ratio_after = proc->mFicHandler.get_fic_decode_ratio_percent();
    f.fibBits.resize(3072);
    bool v[4];
    proc->mFicHandler.get_fib_bits(f.fibBits.data(), v);
    for (int i = 0; i < 4; i++) f.info.fic_valid[i] = v[i] ? 1 : 0;
    framesStarted++;
    start_backends_for_frame(framesStarted); // Backends that shall see their first CIF in the next frame
    return cfg.track_cif ? cifCount : 0; // default: every null symbol is treated as non-TII
  }

  // process_FIB of the stub FIB decoder: count, and (track_cif) follow FIG 0/0's CIF counter the way FibDecoder does
  // (fib_decoder.cpp:59-106 FIG walk, fib_decoder_fig0.cpp:89-113), so that DabProcessor's own TII null-symbol handling runs
  void on_good_fib(const u8 * fib, int) override
  {
    goodFibs++;
    if (!cfg.track_cif) return;
    int done = 0;
    while (done < 30)
    {
      const u8 * d = fib + 8 * done;
      unsigned type = 0, len = 0, ext = 0, hi = 0, lo = 0;
      for (int b = 0; b < 3; b++) type = (type << 1) | (d[b] & 1u);
      for (int b = 3; b < 8; b++) len = (len << 1) | (d[b] & 1u);
      if ((type == 7 && len == 31) || done + (int)len + 1 > 30) break;
      for (int b = 11; b < 16; b++) ext = (ext << 1) | (d[b] & 1u);
      if (type == 0 && ext == 0 && len >= 5)
      {
        for (int b = 35; b < 40; b++) hi = (hi << 1) | (d[b] & 1u);
        for (int b = 40; b < 48; b++) lo = (lo << 1) | (d[b] & 1u);
        cifCount = (int)(hi * 250 + lo);
      }
      done += (int)len + 1;
    }
  }
  void on_msc_frame(int id, const u8 * d, int n) override { auto & v = msc[id]; v.insert(v.end(), d, d + n); }
  void on_lcd(float snr, float mer) override { lastSnr = snr; lastMer = mer; }
};
}

extern "C" void * dabref_chain_run(const float * iq, int64_t n_samples, const dabref_chain_cfg * cfg)
{
  auto * c = new ChainRun;
  c->cfg = *cfg;
  for (int i = 0; i < cfg->n_subch; i++)
  {
    const int32_t * s = cfg->subch + 7 * i;
    c->subch.push_back({ s[0], s[1], s[2], s[3], s[4], s[5], s[6] });
  }
  c->params.threshold = cfg->threshold;
  c->params.tiiFramesToCount = 5;
  c->params.iqBuffer = &c->iqRing;
  c->params.carrBuffer = &c->carrRing;
  c->dev = std::make_unique<MemoryDevice>(reinterpret_cast<const cf32 *>(iq), n_samples);
  ScopedHooks s(c);
  c->proc = std::make_unique<DabProcessor>(&gRadio, c->dev.get(), &c->params);
  c->proc->set_sync_on_strongest_peak(cfg->strongest_peak != 0);
  c->proc->set_dc_avoidance_algorithm(false);
  c->proc->set_dc_and_iq_correction(false, false);
  c->proc->slot_soft_bit_gen_type((ESoftBitType)cfg->soft_bit_type);
  c->proc->set_scan_mode(cfg->scan_mode != 0);
  c->proc->set_tii_processing(false);
  if (cfg->eti_path != nullptr) c->proc->start_eti_generator(QString(cfg->eti_path));
  // start() wipes the service list (dab_processor.cpp:92-95), so Backends for frame 0 are created at the first read.
  c->dev->mOnFirstRead = [c]() { c->start_backends_for_frame(0); };
  const auto t0 = std::chrono::steady_clock::now();
  c->proc->start(); // QThread stub: runs DabProcessor::run() synchronously until the device throws 20
  const auto t1 = std::chrono::steady_clock::now();
  c->seconds = std::chrono::duration<double>(t1 - t0).count();
  if (cfg->eti_path != nullptr) c->proc->stop_eti_generator();
  // A frame cut off by the end of data has no null-symbol record: drop it.
  while (!c->frames.empty() && c->frames.back().fibBits.empty()) c->frames.pop_back();
  return c;
}

extern "C" void dabref_chain_free(void * h)
{
  auto * c = static_cast<ChainRun *>(h);
  ScopedHooks s(c);
  c->proc.reset();
  delete c;
}
extern "C" int dabref_chain_n_frames(void * h) { return (int)static_cast<ChainRun *>(h)->frames.size(); }
extern "C" void dabref_chain_frame_info(void * h, int frame, dabref_frame_info * out) { *out = static_cast<ChainRun *>(h)->frames[frame].info; }
extern "C" void dabref_chain_fib_bits(void * h, int frame, uint8_t out[3072]) { memcpy(out, static_cast<ChainRun *>(h)->frames[frame].fibBits.data(), 3072); }
extern "C" int dabref_chain_soft_bits(void * h, int frame, int16_t * out)
{
  auto & f = static_cast<ChainRun *>(h)->frames[frame];
  if (f.soft.empty()) return 0;
  memcpy(out, f.soft.data(), f.soft.size() * sizeof(i16));
  return 1;
}
extern "C" int dabref_chain_fft(void * h, int frame, float * out)
{
  auto & f = static_cast<ChainRun *>(h)->frames[frame];
  if (f.fft.empty()) return 0;
  memcpy(out, f.fft.data(), f.fft.size() * sizeof(float));
  return 1;
}
extern "C" int dabref_chain_n_good_fibs(void * h) { return static_cast<ChainRun *>(h)->goodFibs; }
extern "C" int64_t dabref_chain_msc_size(void * h, int id)
{
  auto & m = static_cast<ChainRun *>(h)->msc;
  auto it = m.find(id);
  return it == m.end() ? 0 : (int64_t)it->second.size();
}
extern "C" int64_t dabref_chain_msc_copy(void * h, int id, uint8_t * out, int64_t cap)
{
  auto & m = static_cast<ChainRun *>(h)->msc;
  auto it = m.find(id);
  if (it == m.end()) return 0;
  const int64_t n = std::min<int64_t>(cap, (int64_t)it->second.size());
  memcpy(out, it->second.data(), n);
  return n;
}
extern "C" void dabref_chain_counters(void * h, int64_t out[8])
{
  auto * c = static_cast<ChainRun *>(h);
  out[0] = c->dipFoundCount;
  out[1] = c->noDipCount;
  out[2] = c->dev->mSingleReads;
  out[3] = c->dev->mPos;
  out[4] = (int64_t)(c->lastSnr * 1000.0f);
  out[5] = (int64_t)(c->lastMer * 1000.0f);
  out[6] = out[7] = 0;
}
extern "C" double dabref_chain_seconds(void * h) { return static_cast<ChainRun *>(h)->seconds; }
