// Stand-in for the GUI's settings store: the one switch the DAB+ path reads (packet-loss concealment, on by default).
#pragma once
namespace Settings
{
struct BoolValue { bool toBool() const { return true; } };
struct Item { BoolValue read() const { return {}; } };
namespace Config { static const Item cbAudioConcealment; }
}
