// oracle/_ref link-time stand-ins (TEST INFRASTRUCTURE, not product code, not reference code).
//
// The reference's hot-path sources are compiled verbatim from /root/reference; this file supplies the
// symbols they expect from code that is outside the hot path or not installed here:
//   * the moc-less Qt signals (empty bodies; a few record their argument for the harness),
//   * FFTW3f (third-party, un-vendored, call sites dab_processor.cpp:63,201,276,338 and
//     phasereference.cpp:51-52,61,92,105,233,247): unnormalised radix-2 DFT in double, rounded to float,
//   * FibDecoderFactory::create (fib_decoder_if.h:121-125): a recording FIB consumer,
//   * BackendDriver (backend_driver.cpp:38-65): a capture sink instead of the MP2/AAC/data decoders.
#include "harness_internal.h"

#include <cmath>
#include <complex>
#include <vector>

// ---------------------------------------------------------------------------------------------------
// Qt signals declared in reference headers (no moc here, so they need bodies).
// ---------------------------------------------------------------------------------------------------
void DabProcessor::signal_no_dip_sync_found() { if (gHooks) gHooks->noDipCount++; }
void DabProcessor::signal_dip_sync_found() { if (gHooks) gHooks->dipFoundCount++; }
void DabProcessor::signal_show_tii(const std::vector<STiiResult> &) {}
void DabProcessor::signal_show_spectrum(i32) {}
void DabProcessor::signal_show_clock_err(f32) {}
void DabProcessor::signal_set_and_show_freq_corr_rf_Hz(i32) {}
void DabProcessor::signal_show_freq_corr_bb_Hz(i32) {}
void DabProcessor::signal_linear_peak_and_rms_level(f32, f32) {}
void OfdmDecoder::signal_slot_show_iq(i32, f32) {}
void OfdmDecoder::signal_show_lcd_data(const SLcdData & d) { if (gHooks) gHooks->on_lcd(d.SNR, d.MER); }
void PhaseReference::signal_show_correlation(f32, const QVector<i32> &) {}
void SampleReader::signal_show_spectrum(i32) {}
void SampleReader::signal_show_cir(i32) {}
void FicDecoder::signal_fic_status(i32 r, f32 ber) { if (gHooks) gHooks->on_fic_status(r, ber); }
void IFibDecoder::signal_name_of_ensemble(i32, const QString &, const QString &) {}
void IFibDecoder::signal_change_in_configuration() {}
void IFibDecoder::signal_start_announcement(const QString &, i32) {}
void IFibDecoder::signal_stop_announcement(const QString &, i32) {}
void IFibDecoder::signal_fib_time_info(const SUtcTimeSet &) {}
void IFibDecoder::signal_fib_loaded_state(EFibLoadingState) {}

// ---------------------------------------------------------------------------------------------------
// FFTW3f shim: unnormalised DFT, sign as planned, double precision internally.
// ---------------------------------------------------------------------------------------------------
struct fftwf_plan_s
{
  int n;
  int sign;
  fftwf_complex * in;
  fftwf_complex * out;
  std::vector<std::complex<double>> tw;   // e^{sign*j*2*pi*k/n}, k < n/2
  std::vector<int> rev;
  std::vector<std::complex<double>> work;
};

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex * in, fftwf_complex * out, int sign, unsigned)
{
  auto * p = new fftwf_plan_s;
  p->n = n;
  p->sign = sign;
  p->in = in;
  p->out = out;
  p->tw.resize(n / 2);
  for (int k = 0; k < n / 2; k++)
  {
    const double a = (double)sign * 2.0 * M_PI * (double)k / (double)n;
    p->tw[k] = { std::cos(a), std::sin(a) };
  }
  int bits = 0;
  while ((1 << bits) < n) bits++;
  p->rev.resize(n);
  for (int i = 0; i < n; i++)
  {
    int r = 0;
    for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    p->rev[i] = r;
  }
  p->work.resize(n);
  return p;
}

extern "C" void fftwf_execute(const fftwf_plan p)
{
  if (gHooks) gHooks->before_fft(p);
  const int n = p->n;
  auto & w = p->work;
  for (int i = 0; i < n; i++) w[p->rev[i]] = { (double)p->in[i][0], (double)p->in[i][1] };
  for (int len = 2; len <= n; len <<= 1)
  {
    const int half = len / 2;
    const int step = n / len;
    for (int s = 0; s < n; s += len)
    {
      for (int k = 0; k < half; k++)
      {
        const std::complex<double> t = w[s + k + half] * p->tw[k * step];
        w[s + k + half] = w[s + k] - t;
        w[s + k] += t;
      }
    }
  }
  for (int i = 0; i < n; i++)
  {
    p->out[i][0] = (float)w[i].real();
    p->out[i][1] = (float)w[i].imag();
  }
  if (gHooks) gHooks->after_fft(p, (const float *)p->out);
}

extern "C" void fftwf_destroy_plan(fftwf_plan p) { delete p; }

// ---------------------------------------------------------------------------------------------------
// FIB consumer: records CRC-good FIBs, serves a configured sub-channel table and CIF counter.
// ---------------------------------------------------------------------------------------------------
class RecordingFibDecoder : public IFibDecoder
{
public:
  void process_FIB(const std::array<std::byte, cFibSizeVitOut> & iFib, u16 iFicIdx) override
  {
    if (gHooks) gHooks->on_good_fib(reinterpret_cast<const u8 *>(iFib.data()), iFicIdx);
  }
  void connect_channel() override {}
  void disconnect_channel() override {}
  void set_SId_for_fast_audio_selection(u32) override {}
  void get_data_for_audio_service(u32, SAudioData &) const override {}
  void get_data_for_audio_service_addon(u32, SAudioDataAddOns &) const override {}
  void get_data_for_packet_service(u32, std::vector<SPacketData> &) const override {}
  std::vector<SServiceId> get_service_list() const override { return {}; }
  const QString & get_service_label_from_SId_SCIdS(u32, i32) const override { return mEmpty; }
  void get_SId_SCIdS_from_service_label(const QString &, u32 & oSId, i32 & oSCIdS) const override { oSId = 0; oSCIdS = 0; }
  u8 get_ecc() const override { return 0; }
  i32 get_EId() const override { return 0; }
  QString get_ensemble_name() const override { return {}; }
  std::vector<i8> get_sub_channel_id_list() const override
  {
    std::vector<i8> v;
    if (gHooks) for (const auto & s : gHooks->subch) v.push_back((i8)s.subChId);
    return v;
  }
  void get_sub_channel_info(SChannelData * d, i32 iSubChId) const override
  {
    d->in_use = false;
    if (!gHooks) return;
    for (const auto & s : gHooks->subch)
    {
      if (s.subChId == iSubChId)
      {
        d->in_use = true;
        d->id = (i16)s.subChId;
        d->start_cu = (i16)s.startCU;
        d->uepFlag = (u8)s.shortForm;
        d->protlev = (i16)s.protLevel;
        d->size = (i16)s.sizeCU;
        d->bitrate = (i16)s.bitRate;
        d->ASCTy = 077;
      }
    }
  }
  // Called once per frame by DabProcessor::_process_null_symbol (dab_processor.cpp:274): frame boundary tap.
  i32 get_cif_count() const override { return gHooks ? gHooks->on_get_cif_count() : 0; }
  void get_cif_count(i16 * h, i16 * l) const override { *h = 0; *l = 0; }
  u32 get_mod_julian_date() const override { return 0; }
  QStringList get_fib_content_str_list(i32 & oNumCols) const override { oNumCols = 0; return {}; }

private:
  QString mEmpty;
};

std::unique_ptr<IFibDecoder> FibDecoderFactory::create(DabRadio *) { return std::make_unique<RecordingFibDecoder>(); }

// ---------------------------------------------------------------------------------------------------
// BackendDriver: capture sink.
// ---------------------------------------------------------------------------------------------------
class CaptureProcessor : public FrameProcessor
{
public:
  explicit CaptureProcessor(int iSubChId) : mSubChId(iSubChId) {}
  void add_to_frame(const std::vector<u8> & v) override { if (gHooks) gHooks->on_msc_frame(mSubChId, v.data(), (int)v.size()); }
private:
  const int mSubChId;
};

BackendDriver::BackendDriver(DabRadio *, const SDescriptorType * ipDT, RingBuffer<i16> *, RingBuffer<u8> *, RingBuffer<u8> *)
{
  mpFrameProcessor = std::make_unique<CaptureProcessor>(ipDT->SubChId);
}

void BackendDriver::add_to_frame(const std::vector<u8> & iData) const { mpFrameProcessor->add_to_frame(iData); }
