// file_probe.cu — what the reference's file readers learn from a file before they start reading samples (host code):
//   * XmlDescriptor (devices/filereaders/xml_filereader/xml_descriptor.cpp:100-242): the <SDR> header of a UFF / "xml"
//     recording (sample rate, container, bits, byte order, channel order, data blocks) and XmlFileReader's start offset
//     (xml_filereader.cpp:103-126);
//   * WavFileHandler (wav_files/wavfiles.cpp:61-97), which asks libsndfile for the sample rate, the channel count and
//     the sample format of a RIFF/WAVE file (.wav / .sdr recordings);
//   * RawReader (raw_files/raw_reader.cpp:66-70): no header, unsigned 8-bit IQ at 2.048 MS/s.
// The reference parses the XML with QDomDocument and the WAVE file with libsndfile; neither library is part of the
// reference tree or this image, so this is a restatement of the fields the readers use, not of those libraries.
#include "../../include/dabstar_b200.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace
{
struct Tag
{
  std::string name;                                       // "Samplerate", "/Sample", ...
  std::vector<std::pair<std::string, std::string>> attr;  // in document order
  bool self_closing = false;
  const std::string * get(const char * key) const
  {
    for (auto & a : attr) if (a.first == key) return &a.second;
    return nullptr;
  }
  std::string value(const char * key, const char * dflt) const { const std::string * v = get(key); return v ? *v : std::string(dflt); }
};

// Splits the header text into tags; comments, processing instructions and character data are skipped.
std::vector<Tag> scan_tags(const char * p, size_t n)
{
  std::vector<Tag> tags;
  size_t i = 0;
  while (i < n)
  {
    if (p[i] != '<') { i++; continue; }
    if (i + 3 < n && p[i + 1] == '!' && p[i + 2] == '-' && p[i + 3] == '-')
    {
      const char * e = (const char *)memmem(p + i, n - i, "-->", 3);
      if (!e) break;
      i = (size_t)(e - p) + 3;
      continue;
    }
    if (i + 1 < n && (p[i + 1] == '?' || p[i + 1] == '!'))
    {
      const char * e = (const char *)memchr(p + i, '>', n - i);
      if (!e) break;
      i = (size_t)(e - p) + 1;
      continue;
    }
    size_t j = i + 1;
    Tag t;
    while (j < n && p[j] != '>' && p[j] != '/' && !isspace((unsigned char)p[j])) t.name.push_back(p[j++]);
    if (t.name.empty() && j < n && p[j] == '/')
    {
      t.name = "/";
      j++;
      while (j < n && p[j] != '>' && !isspace((unsigned char)p[j])) t.name.push_back(p[j++]);
    }
    while (j < n && p[j] != '>')
    {
      while (j < n && isspace((unsigned char)p[j])) j++;
      if (j < n && p[j] == '/') { t.self_closing = true; j++; continue; }
      if (j >= n || p[j] == '>') break;
      std::string key, val;
      while (j < n && p[j] != '=' && p[j] != '>' && !isspace((unsigned char)p[j])) key.push_back(p[j++]);
      while (j < n && isspace((unsigned char)p[j])) j++;
      if (j < n && p[j] == '=')
      {
        j++;
        while (j < n && isspace((unsigned char)p[j])) j++;
        if (j < n && (p[j] == '"' || p[j] == '\''))
        {
          const char q = p[j++];
          while (j < n && p[j] != q) val.push_back(p[j++]);
          if (j < n) j++;
        }
      }
      if (!key.empty()) t.attr.emplace_back(key, val);
    }
    if (j >= n) break;
    tags.push_back(std::move(t));
    i = j + 1;
  }
  return tags;
}

std::string upper(std::string s)
{
  for (auto & c : s) c = (char)toupper((unsigned char)c);
  return s;
}

int container_of(const std::string & c)
{
  static const char * const names[6] = { "int8", "uint8", "int16", "int24", "int32", "float32" };
  for (int k = 0; k < 6; k++) if (c == names[k]) return k;
  return -1;
}

int probe_xml(const uint8_t * head, int64_t head_bytes, int64_t file_bytes, dabstar_file_info * out)
{
  // XmlDescriptor reads characters until it has seen 500 zero bytes in a row (xml_descriptor.cpp:115-129)
  int64_t end = 0, zeros = 0;
  while (end < head_bytes && zeros < 500) { zeros = head[end] == 0 ? zeros + 1 : 0; end++; }
  const size_t text_len = strnlen((const char *)head, (size_t)end);
  const std::vector<Tag> tags = scan_tags((const char *)head, text_len);

  // defaults of the descriptor (xml_descriptor.cpp:108-113); its container default "i16" matches no reader branch
  int sample_rate = 2048000, bits = 16, n_channels = 2;
  std::string container = "i16", ordering = "MSB", iq_order = "IQ";
  int64_t n_elements = 0, samples_to_read = 0;
  int n_blocks = 0, frequency = 0;
  bool in_sample = false, in_channels = false, in_blocks = false, in_block = false;
  int channel_order = 0;
  std::vector<std::pair<int64_t, std::string>> blocks; // (Count, unit)
  for (const Tag & t : tags)
  {
    if (t.name == "Sample") in_sample = !t.self_closing;
    else if (t.name == "/Sample") in_sample = false;
    else if (t.name == "Datablocks") in_blocks = !t.self_closing;
    else if (t.name == "/Datablocks") in_blocks = false;
    else if (in_sample && t.name == "Samplerate")
    {
      const std::string hz = t.value("Unit", "Hz");
      const int factor = hz == "Hz" ? 1 : ((hz == "KHz" || hz == "Khz") ? 1000 : 1000000); // xml_descriptor.cpp:165
      sample_rate = atoi(t.value("Value", "2048000").c_str()) * factor;
    }
    else if (in_sample && t.name == "Channels")
    {
      n_channels = atoi(t.value("Amount", "2").c_str());
      bits = atoi(t.value("Bits", "8").c_str());
      container = t.value("Container", "u8");
      ordering = t.value("Ordering", "N/A");
      in_channels = !t.self_closing;
      channel_order = 0;
    }
    else if (t.name == "/Channels") in_channels = false;
    else if (in_channels && t.name == "Channel")
    {
      // addChannelOrder (xml_descriptor.cpp:56-74)
      const std::string v = t.value("Value", "I");
      if (channel_order == 0) iq_order = v == "I" ? "I_ONLY" : "Q_ONLY";
      else if (channel_order == 1 && iq_order == "I_ONLY" && v == "Q") iq_order = "IQ";
      else if (channel_order == 1 && iq_order == "Q_ONLY" && v == "I") iq_order = "QI";
      channel_order++;
    }
    else if (in_blocks && t.name == "Datablock")
    {
      blocks.emplace_back(atoll(t.value("Count", "100").c_str()), t.value("Channel", "Channel")); // (the unit is read from an attribute named "Channel", :213)
      in_block = !t.self_closing;
      n_blocks++;
    }
    else if (t.name == "/Datablock") in_block = false;
    else if (in_block && t.name == "Frequency" && n_blocks == 1)
    {
      const std::string u = upper(t.value("Unit", "Hz"));
      const int v = atoi(t.value("Value", "200").c_str());
      frequency = u == "HZ" ? v : (u == "KHZ" ? v * 1000 : v * 1000000);
    }
  }
  if (n_blocks <= 0) return DABSTAR_E_INVALID; // *ok = nrBlocks > 0
  const bool pairs = iq_order == "IQ" || iq_order == "QI";
  for (auto & b : blocks)
  {
    n_elements += b.first;
    samples_to_read += (b.second == "Channel" && pairs) ? b.first / 2 : b.first; // compute_nrSamples, xml_filereader.cpp:313-335
  }
  const int cont = container_of(container);
  if (cont < 0) return DABSTAR_E_INVALID;
  const int bytecount = cont <= 1 ? 1 : (cont == 2 ? 2 : (cont == 3 ? 3 : 4));
  int64_t start = file_bytes - n_elements * bytecount; // xml_filereader.cpp:124-126
  if (start < 2048 || start > 1000000) start = 2048;

  memset(out, 0, sizeof(*out));
  out->kind = DABSTAR_FILE_XML;
  out->reader = DABSTAR_READER_XML;
  out->sample_rate = sample_rate;
  out->n_channels = n_channels;
  out->fmt.container = cont;
  out->fmt.bits_per_channel = bits;
  out->fmt.msb_first = ordering == "MSB" ? 1 : 0;
  // XmlReader::run compares with "I_Only" while the descriptor writes "I_ONLY": both single-channel layouts end in
  // readElements_Q (xml_reader.cpp:153-168)
  out->fmt.iq_order = iq_order == "IQ" ? DABSTAR_ORDER_IQ : (iq_order == "QI" ? DABSTAR_ORDER_QI : DABSTAR_ORDER_Q_ONLY);
  out->data_offset = start;
  const int elem = bytecount * (pairs ? 2 : 1);
  const int64_t room = file_bytes > start ? (file_bytes - start) / elem : 0;
  out->n_samples = samples_to_read < room ? samples_to_read : room;
  out->frequency_hz = frequency;
  return 0;
}

uint32_t le32(const uint8_t * p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const uint8_t * p) { return (uint16_t)(p[0] | (p[1] << 8)); }

int probe_wav(const uint8_t * head, int64_t head_bytes, int64_t file_bytes, dabstar_file_info * out)
{
  if (head_bytes < 12 || memcmp(head + 8, "WAVE", 4) != 0) return DABSTAR_E_INVALID;
  int64_t pos = 12;
  int format_tag = 0, channels = 0, bits = 0, rate = 0;
  bool have_fmt = false;
  while (pos + 8 <= head_bytes)
  {
    const uint8_t * c = head + pos;
    const uint32_t size = le32(c + 4);
    if (memcmp(c, "fmt ", 4) == 0)
    {
      if (size < 16 || pos + 8 + 16 > head_bytes) return DABSTAR_E_INVALID;
      format_tag = le16(c + 8);
      channels = le16(c + 10);
      rate = (int)le32(c + 12);
      bits = le16(c + 22);
      if (format_tag == 0xFFFE) // WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the plain tag
      {
        if (size < 40 || pos + 8 + 26 > head_bytes) return DABSTAR_E_INVALID;
        format_tag = le16(c + 8 + 24);
      }
      have_fmt = true;
    }
    else if (memcmp(c, "data", 4) == 0)
    {
      if (!have_fmt) return DABSTAR_E_INVALID;
      // WavFileHandler's checks (wavfiles.cpp:71-97): 1.536 .. 3.0 MS/s, two channels, PCM 8/16/24/32 or float
      if (rate < 1536000 || rate > 3000000 || channels != 2) return DABSTAR_E_INVALID;
      int cont;
      if (format_tag == 1) cont = bits == 8 ? DABSTAR_CONTAINER_UINT8_PCM : (bits == 16 ? DABSTAR_CONTAINER_INT16 : (bits == 24 ? DABSTAR_CONTAINER_INT24 : (bits == 32 ? DABSTAR_CONTAINER_INT32_PCM : -1)));
      else if (format_tag == 3 && bits == 32) cont = DABSTAR_CONTAINER_FLOAT32;
      else cont = -1;
      if (cont < 0) return DABSTAR_E_INVALID;
      memset(out, 0, sizeof(*out));
      out->kind = DABSTAR_FILE_WAV;
      out->reader = DABSTAR_READER_WAV;
      out->sample_rate = rate;
      out->n_channels = channels;
      out->fmt.container = cont;
      out->fmt.bits_per_channel = 0; // libsndfile scales by the container width
      out->fmt.msb_first = 0;
      out->fmt.iq_order = DABSTAR_ORDER_IQ;
      out->data_offset = pos + 8;
      int64_t bytes = size;
      if (bytes == 0xFFFFFFFFll || out->data_offset + bytes > file_bytes) bytes = file_bytes - out->data_offset; // streamed or truncated files
      out->n_samples = bytes / (2 * (bits / 8));
      return 0;
    }
    pos += 8 + (int64_t)size + (size & 1);
  }
  return DABSTAR_E_INVALID;
}
} // namespace

extern "C" int dabstar_file_probe(const uint8_t * head, int64_t head_bytes, int64_t file_bytes, dabstar_file_info * out)
{
  if (!head || !out || head_bytes < 0 || file_bytes < head_bytes) return DABSTAR_E_INVALID;
  if (head_bytes >= 12 && (memcmp(head, "RIFF", 4) == 0 || memcmp(head, "RF64", 4) == 0)) return probe_wav(head, head_bytes, file_bytes, out);
  // a UFF header is text: "<?xml" or "<SDR" after optional white space
  int64_t i = 0;
  while (i < head_bytes && i < 64 && isspace(head[i])) i++;
  if (i + 4 < head_bytes && head[i] == '<' && (memcmp(head + i, "<?xml", 5) == 0 || memcmp(head + i, "<SDR", 4) == 0)) return probe_xml(head, head_bytes, file_bytes, out);
  // RawReader: unsigned bytes, I then Q, 2.048 MS/s (raw_reader.cpp:66-70)
  memset(out, 0, sizeof(*out));
  out->kind = DABSTAR_FILE_RAW;
  out->reader = DABSTAR_READER_XML;
  out->sample_rate = 2048000;
  out->n_channels = 2;
  out->fmt.container = DABSTAR_CONTAINER_UINT8;
  out->fmt.bits_per_channel = 8;
  out->fmt.iq_order = DABSTAR_ORDER_IQ;
  out->data_offset = 0;
  out->n_samples = file_bytes / 2;
  return 0;
}
