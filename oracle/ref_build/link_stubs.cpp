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
// backend/charsets.cpp needs QChar tables; labels are not part of what is pinned: bytes pass through
QString to_QString_using_charset(const QByteArray & b, ECharacterSet) { return QString(std::string(b)); }
QString to_QString_using_charset(const char * p, ECharacterSet, i32 size) { return QString::fromUtf8(p, size); }
// the file readers: XmlDescriptor's XML parsing (QDomDocument) is not built, the harness fills the fields; signals of the threads
XmlDescriptor::XmlDescriptor(FILE *, bool * ok) { sampleRate = 0; nrChannels = 2; bitsperChannel = 0; nrBlocks = 0; if (ok) *ok = true; }
void XmlReader::signal_set_progress(i64, i64) {}
void XmlReader::signal_file_looped() {}
void WavReader::signal_set_progress(i32, f32) {}
void WavReader::signal_file_looped() {}
// Mp4Processor's signals (moc would generate them): GUI counters only
void Mp4Processor::signal_show_frame_errors(i32) {}
void Mp4Processor::signal_show_rs_errors(i32) {}
void Mp4Processor::signal_show_aac_errors(i32) {}
void Mp4Processor::signal_is_stereo(bool) {}
void Mp4Processor::signal_new_aac_frame() {}
void Mp4Processor::signal_show_rs_corrections(i32, i32) {}

// ---------------------------------------------------------------------------------------------------
// FFTW3f shim: unnormalised DFT, sign as planned.
// Parity build (default): radix-2 in double precision, rounded to float (the golden vectors were made with it).
// Timing build (-DDABREF_FAST_FFT, libdabref_fast.so): single-precision Stockham autosort radix-4 (+ one radix-2 pass for
// n = 2^odd) with precomputed twiddles, the inner loops written so that gcc vectorises them: what stands in for FFTW's
// speed when the reference's CPU chain is TIMED (bench.py); FFTW itself is third party and not installed.
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
#ifdef DABREF_FAST_FFT
  std::vector<float> twr, twi;            // e^{sign*j*2*pi*k/n}, k < n (split re / im)
  std::vector<float> ar, ai, br, bi;      // ping-pong work arrays (split re / im)
#endif
};

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex * in, fftwf_complex * out, int sign, unsigned)
{
  auto * p = new fftwf_plan_s;
  p->n = n;
  p->sign = sign;
  p->in = in;
  p->out = out;
#ifdef DABREF_FAST_FFT
  p->twr.resize(n); p->twi.resize(n);
  for (int k = 0; k < n; k++)
  {
    const double a = (double)sign * 2.0 * M_PI * (double)k / (double)n;
    p->twr[k] = (float)std::cos(a);
    p->twi[k] = (float)std::sin(a);
  }
  p->ar.resize(n); p->ai.resize(n); p->br.resize(n); p->bi.resize(n);
#else
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
#endif
  return p;
}

#ifdef DABREF_FAST_FFT
// One Stockham pass of radix R over x (length n = R * l * m): y[(R j + q) m + k] = sum_r w^{j l_stride ...}; standard autosort form:
// for j < l, k < m:  a_r = x[(j + r l) m + k];  y[(R j + q) m + k] = (sum_r a_r W_R^{r q}) * W_n^{q j m ... } (twiddle applied to inputs).
static void stockham_radix4(int n, int l, int m, int sign, const float * xr, const float * xi, float * yr, float * yi, const float * twr, const float * twi)
{
  // x viewed as [4][l][m], y as [l][4][m]; twiddle w = W_n^{j m}
  for (int j = 0; j < l; j++)
  {
    const int t1 = (j * m) % n, t2 = (2 * j * m) % n, t3 = (3 * j * m) % n;
    const float w1r = twr[t1], w1i = twi[t1], w2r = twr[t2], w2i = twi[t2], w3r = twr[t3], w3i = twi[t3];
    const float * x0r = xr + (size_t)j * m, * x0i = xi + (size_t)j * m;
    const float * x1r = x0r + (size_t)l * m, * x1i = x0i + (size_t)l * m;
    const float * x2r = x1r + (size_t)l * m, * x2i = x1i + (size_t)l * m;
    const float * x3r = x2r + (size_t)l * m, * x3i = x2i + (size_t)l * m;
    float * y0r = yr + (size_t)4 * j * m, * y0i = yi + (size_t)4 * j * m;
    float * y1r = y0r + m, * y1i = y0i + m, * y2r = y1r + m, * y2i = y1i + m, * y3r = y2r + m, * y3i = y2i + m;
    const float sg = (float)sign; // -j for the forward transform (sign = -1): multiply by sign * j
    for (int k = 0; k < m; k++)
    {
      const float ar = x0r[k], ai = x0i[k], br_ = x1r[k], bi_ = x1i[k], cr = x2r[k], ci = x2i[k], dr = x3r[k], di = x3i[k];
      const float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
      const float s2r = br_ + dr, s2i = bi_ + di, s3r = br_ - dr, s3i = bi_ - di;
      // sign * j * s3
      const float j3r = -sg * s3i, j3i = sg * s3r;
      const float u0r = s0r + s2r, u0i = s0i + s2i;
      const float u1r = s1r + j3r, u1i = s1i + j3i;
      const float u2r = s0r - s2r, u2i = s0i - s2i;
      const float u3r = s1r - j3r, u3i = s1i - j3i;
      y0r[k] = u0r; y0i[k] = u0i;
      y1r[k] = u1r * w1r - u1i * w1i; y1i[k] = u1r * w1i + u1i * w1r;
      y2r[k] = u2r * w2r - u2i * w2i; y2i[k] = u2r * w2i + u2i * w2r;
      y3r[k] = u3r * w3r - u3i * w3i; y3i[k] = u3r * w3i + u3i * w3r;
    }
  }
}
static void stockham_radix2(int n, int l, int m, const float * xr, const float * xi, float * yr, float * yi, const float * twr, const float * twi)
{
  for (int j = 0; j < l; j++)
  {
    const int t = (j * m) % n;
    const float wr = twr[t], wi = twi[t];
    const float * x0r = xr + (size_t)j * m, * x0i = xi + (size_t)j * m, * x1r = x0r + (size_t)l * m, * x1i = x0i + (size_t)l * m;
    float * y0r = yr + (size_t)2 * j * m, * y0i = yi + (size_t)2 * j * m, * y1r = y0r + m, * y1i = y0i + m;
    for (int k = 0; k < m; k++)
    {
      const float ar = x0r[k], ai = x0i[k], br_ = x1r[k], bi_ = x1i[k];
      y0r[k] = ar + br_; y0i[k] = ai + bi_;
      const float dr = ar - br_, di = ai - bi_;
      y1r[k] = dr * wr - di * wi; y1i[k] = dr * wi + di * wr;
    }
  }
}
#endif

extern "C" void fftwf_execute(const fftwf_plan p)
{
  if (gHooks) gHooks->before_fft(p);
  const int n = p->n;
#ifdef DABREF_FAST_FFT
  float * xr = p->ar.data(), * xi = p->ai.data(), * yr = p->br.data(), * yi = p->bi.data();
  for (int i = 0; i < n; i++) { xr[i] = p->in[i][0]; xi[i] = p->in[i][1]; }
  // decimation in frequency, Stockham autosort: x is [R][l][m] with l * m * R = n; m grows by R every pass
  int l = n, m = 1;
  while (l > 1)
  {
    if (l % 4 == 0) { l /= 4; stockham_radix4(n, l, m, p->sign, xr, xi, yr, yi, p->twr.data(), p->twi.data()); m *= 4; }
    else { l /= 2; stockham_radix2(n, l, m, xr, xi, yr, yi, p->twr.data(), p->twi.data()); m *= 2; }
    std::swap(xr, yr); std::swap(xi, yi);
  }
  for (int i = 0; i < n; i++) { p->out[i][0] = xr[i]; p->out[i][1] = xi[i]; }
#else
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
#endif
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
