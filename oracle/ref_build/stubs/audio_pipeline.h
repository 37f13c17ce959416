// Stand-in for the audio output pipeline (out of scope): only the slot Mp4Processor connects to.
#pragma once
#include <QObject>
class AudioPipeline : public QObject
{
public:
  void slot_new_aac_mp2_frame() {}
};
