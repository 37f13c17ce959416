// TII detection is diagnostic only (out of scope): no-op stand-in.
#pragma once
#include "dab_constants.h"
#include <vector>
struct STiiResult { int mainId = 0; int subId = 0; };
class TiiDetector
{
public:
  void reset() {}
  void set_detect_collisions(bool) {}
  void set_subid_for_collision_search(u8) {}
  void add_to_tii_buffer(const TArrayTu &) {}
  std::vector<STiiResult> process_tii_data(i16) { return {}; }
};
