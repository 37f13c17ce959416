// Shared between harness.cpp and link_stubs.cpp (oracle/_ref build only; test infrastructure).
#pragma once
// std and stub headers first: the access hack below must not reach into libstdc++.
#include <algorithm>
#include <array>
#include <atomic>
#include <complex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <QObject>
#include <QThread>
#include <QMutex>
#include <QSharedPointer>
#include <QStringList>
#include <QDebug>
#include <fftw3.h>
#include <sndfile.h>
#define private public   // taps read a few private members of the reference classes (layout is unaffected)
#define protected public
#include "dab_processor.h"
#include "eti_generator.h"
#include "backend.h"
#include "backend_driver.h"
#include "eep_protection.h"
#include "uep_protection.h"
#include "crc.h"
#include "dabradio.h"
#include "process_params.h"
#include "mp4processor.h"
#include "fib_decoder.h"
#include "charsets.h"
#include "xml_reader.h"
#include "xml_descriptor.h"
#include "xml_filereader.h"
#include "wav_reader.h"
#include "wavfiles.h"
#undef private
#undef protected
#include <vector>

struct RefSubCh
{
  int subChId, startCU, sizeCU, shortForm, protLevel, bitRate, startFrame;
};

struct Hooks
{
  std::vector<RefSubCh> subch;
  int noDipCount = 0;
  int dipFoundCount = 0;
  virtual ~Hooks() = default;
  virtual void before_fft(fftwf_plan) {}
  virtual void after_fft(fftwf_plan, const float *) {}
  virtual void on_good_fib(const u8 *, int) {}
  virtual int on_get_cif_count() { return 0; }
  virtual void on_msc_frame(int, const u8 *, int) {}
  virtual void on_fic_status(int, float) {}
  virtual void on_lcd(float, float) {}
};

extern Hooks * gHooks;
