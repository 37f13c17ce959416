// Stand-in for backend/data/pad_handler.h: PAD / MOT handling is out of scope; Mp4Processor only forwards to it.
#pragma once
#include "glob_data_types.h"
class DabRadio;
class PadHandler
{
public:
  explicit PadHandler(DabRadio *) {}
  void process_PAD(const u8 *, i16, u8, u8) {}
};
