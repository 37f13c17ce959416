// Stand-in for backend/audio/faad_decoder.h (libfaad is third party and not installed): the same interface, but instead of
// decoding AAC it records what Mp4Processor hands to the decoder. Test infrastructure only (oracle/_ref build).
#pragma once
#include <QObject>
#include <vector>
#include "glob_data_types.h"
#include "ringbuffer.h"
class DabRadio;
struct SStreamParms
{
  i32 dacRate;
  i32 sbrFlag;
  i32 psFlag;
  i32 aacChannelMode;
  i32 mpegSurround;
  i32 CoreChConfig;
  i32 CoreSrIndex;
  i32 ExtensionSrIndex;
};
struct DabrefAacEvent
{
  int kind;                 // 1: access unit handed to convert_mp4_to_pcm, 0: conceal_lost_frame
  int value;                // kind 0: samples to conceal; kind 1: dacRate | sbrFlag << 1 | aacChannelMode << 2 | psFlag << 3 | mpegSurround << 4
  std::vector<u8> bytes;    // kind 1: the access unit (without its CRC)
};
extern std::vector<DabrefAacEvent> * gDabrefAacSink; // where the stand-in files its calls (set by the harness)
class FaadDecoder : public QObject
{
public:
  FaadDecoder(DabRadio *, RingBuffer<i16> *) {}
  i16 convert_mp4_to_pcm(const SStreamParms * sp, const u8 * buf, i16 len)
  {
    if (gDabrefAacSink)
      gDabrefAacSink->push_back({ 1, sp->dacRate | (sp->sbrFlag << 1) | (sp->aacChannelMode << 2) | (sp->psFlag << 3) | (sp->mpegSurround << 4), std::vector<u8>(buf, buf + len) });
    return 1; // "samples decoded": any positive value keeps Mp4Processor on its good path
  }
  void conceal_lost_frame(i32 n)
  {
    if (gDabrefAacSink) gDabrefAacSink->push_back({ 0, n, {} });
  }
};
