// Stand-in for devices/filereaders/wav_files/wavfiles.h (a Qt widget): what WavReader touches of its parent.
#pragma once
#include <QObject>
#include "glob_data_types.h"
#include "xml_filereader.h"
class WavFileHandler : public QObject
{
public:
  DabrefCheckBox box;
  DabrefCheckBox * cbLoopFile = &box;
  void slot_set_progress(i32, f32) const {}
  void signal_file_looped() {}
};
