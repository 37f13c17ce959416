// Stand-in for the GUI main window: only the slot names the hot path connects to.
#pragma once
#include <QObject>
#include <QVector>
#include <QTimer>
#include <vector>
#include "dab_constants.h"
#include "glob_enums.h"
struct STiiResult;
class DabRadio : public QObject
{
public:
  void slot_show_spectrum(int) {}
  void slot_show_cir(int) {}
  void slot_show_iq(int, float) {}
  void slot_show_lcd_data(...) {}
  void slot_show_correlation(float, const QVector<int> &) {}
  void slot_show_fic_status(int, float) {}
  void slot_show_tii(const std::vector<STiiResult> &) {}
  void slot_show_clock_error(float) {}
  void slot_set_and_show_freq_corr_rf_Hz(int) {}
  void slot_show_freq_corr_bb_Hz(int) {}
  void slot_show_digital_peak_and_rms_level(float, float) {}
  // Mp4Processor (DAB+ outer code)
  void slot_show_frame_errors(int) {}
  void slot_show_rs_errors(int) {}
  void slot_show_aac_errors(int) {}
  void slot_set_stereo(bool) {}
  void slot_show_rs_corrections(int, int) {}
  class AudioPipeline * get_audio_pipeline() { return nullptr; }
  // FibDecoder
  void slot_name_of_ensemble(...) {}
  void slot_fib_time(...) {}
  void slot_change_in_configuration(...) {}
  void slot_start_announcement(...) {}
  void slot_stop_announcement(...) {}
  void slot_fib_loaded_state(...) {}
};
