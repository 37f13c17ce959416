// Stand-in for devices/filereaders/xml_filereader/xml_filereader.h (a Qt widget): what XmlReader touches of its parent.
#pragma once
#include <QObject>
#include "glob_data_types.h"
struct DabrefCheckBox { bool isChecked() const { return false; } }; // "loop file" unchecked: the reader stops at the end
class XmlFileReader : public QObject
{
public:
  DabrefCheckBox box;
  DabrefCheckBox * cbLoopFile = &box;
  i64 samplesToRead = 0;
  i64 mFileLength = 0;
  void slot_set_progress(i64, i64) {}
  void signal_file_looped() {}
};
