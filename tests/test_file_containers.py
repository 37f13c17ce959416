"""Next row f4: input containers. CPU part: header inspection (XmlDescriptor / WavFileHandler / RawReader, restated in
csrc/file_probe.cu because the reference parses with QDomDocument and libsndfile) on files built here, and the oracle's
restatement of the readers' linear-interpolation resampling against an independent numpy statement. GPU part: the
resampling kernel and the PCM conversions bit exact against the oracle, and a 2.5 MS/s WAV recording decoded through
probe -> convert -> resample -> DabProcessor equals the CPU chain on the oracle-resampled samples.
Parity is unpinned at this boundary (XmlReader / WavReader need Qt and libsndfile): the tests follow the source text."""
import struct

import numpy as np
import pytest

from dabstar_b200 import api, synth


def wav_bytes(samples: np.ndarray, rate: int, fmt_tag: int, bits: int, channels: int = 2, extensible: bool = False, junk: bool = False) -> bytes:
    body = samples.tobytes()
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits, 22, bits, 3, fmt_tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    else:
        fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    chunks = b""
    if junk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"  # odd-sized chunk is padded to even
    chunks += b"fmt " + struct.pack("<I", len(fmt)) + fmt
    chunks += b"data" + struct.pack("<I", len(body)) + body
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def xml_bytes(body: bytes, rate_value: str, rate_unit: str, container: str, bits: int, ordering: str, channels=("I", "Q"), count=None, header_room=2048,
              freq=("kHz", 227360)) -> bytes:
    n_elem = len(body) // {"int8": 1, "uint8": 1, "int16": 2, "int24": 3, "int32": 4, "float32": 4}[container]
    ch = "".join(f'<Channel Value="{c}"/>' for c in channels)
    text = ('<?xml version="1.0" encoding="utf-8"?>\n<SDR>\n <Recorder Name="x" Version="1"/>\n <!-- a comment <Sample> -->\n <Device Name="dev" Model="m"/>\n'
            ' <Time Unit="UTC" Value="2024-01-01 00:00:00"/>\n <Sample>\n'
            f'  <Samplerate Unit="{rate_unit}" Value="{rate_value}"/>\n  <Channels Bits="{bits}" Container="{container}" Ordering="{ordering}">{ch}</Channels>\n </Sample>\n'
            f' <Datablocks>\n  <Datablock Count="{n_elem if count is None else count}" Number="1" Unit="Channel">\n   <Frequency Value="{freq[1]}" Unit="{freq[0]}"/>\n'
            '   <Modulation Value="DAB"/>\n  </Datablock>\n </Datablocks>\n</SDR>\n').encode()
    assert len(text) + 500 < header_room
    return text + bytes(header_room - len(text)) + body


# ------------------------------------------------------------------------------------------------ header inspection (CPU)
@pytest.mark.parametrize("tag,bits,dtype,container", [(1, 16, np.int16, "int16"), (1, 8, np.uint8, "uint8_pcm"), (1, 32, np.int32, "int32_pcm"), (3, 32, np.float32, "float32")])
@pytest.mark.parametrize("extensible,junk", [(False, False), (True, True)])
def test_probe_wav(tag, bits, dtype, container, extensible, junk):
    x = (np.arange(2 * 1000) % 97).astype(dtype)
    f = wav_bytes(x, 2500000, tag, bits, extensible=extensible, junk=junk)
    fi = api.file_probe(f)
    assert (fi.kind, fi.reader, fi.sample_rate, fi.n_channels) == ("wav", api.READER_WAV, 2500000, 2)
    assert (fi.fmt.container, fi.fmt.byte_order, fi.fmt.iq_order) == (container, "LSB", "IQ")
    assert fi.n_samples == 1000 and f[fi.data_offset:fi.data_offset + 4] == x.tobytes()[:4]


def test_probe_wav_24bit_and_truncated_and_rejected():
    body = np.zeros(6 * 300, np.uint8)
    fi = api.file_probe(wav_bytes(body, 2048000, 1, 24))
    assert fi.fmt.container == "int24" and fi.n_samples == 300
    # a data chunk that claims more than the file holds (recording cut short, or a streamed 0xFFFFFFFF size)
    f = bytearray(wav_bytes(np.zeros(2 * 500, np.int16), 2048000, 1, 16))
    i = f.index(b"data")
    f[i + 4:i + 8] = struct.pack("<I", 0xFFFFFFFF)
    assert api.file_probe(bytes(f)).n_samples == 500
    # WavFileHandler refuses mono files and rates outside 1.536 .. 3.0 MS/s (wavfiles.cpp:71)
    for bad in (wav_bytes(np.zeros(100, np.int16), 2048000, 1, 16, channels=1), wav_bytes(np.zeros(100, np.int16), 1000000, 1, 16),
                wav_bytes(np.zeros(100, np.int16), 2048000, 6, 16)):
        with pytest.raises(api.DabstarError):
            api.file_probe(bad)


def test_probe_xml():
    body = np.arange(4 * 6000, dtype=np.uint8).tobytes()
    fi = api.file_probe(xml_bytes(body, "2048", "KHz", "int16", 12, "LSB"))
    assert (fi.kind, fi.reader, fi.sample_rate, fi.frequency_hz) == ("xml", api.READER_XML, 2048000, 227360000)
    assert (fi.fmt.container, fi.fmt.bits_per_channel, fi.fmt.byte_order, fi.fmt.iq_order) == ("int16", 12, "LSB", "IQ")
    assert fi.data_offset == 2048 and fi.n_samples == 6000
    # start = file length - elements * width when that lands in [2048, 1e6] (xml_filereader.cpp:124-126)
    fi = api.file_probe(xml_bytes(body, "2500000", "Hz", "uint8", 8, "N/A", channels=("Q", "I"), header_room=5000))
    assert fi.data_offset == 5000 and fi.fmt.iq_order == "QI" and fi.fmt.container == "uint8" and fi.n_samples == len(body) // 2 and fi.sample_rate == 2500000
    # a Count smaller than the file: the start offset falls outside the window and the reader starts at 2048
    fi = api.file_probe(xml_bytes(bytes(2_000_000), "2", "MHz", "int8", 8, "MSB", count=1000))
    assert fi.data_offset == 2048 and fi.n_samples == 500 and fi.sample_rate == 2000000
    # single-channel recordings: the descriptor writes "I_ONLY", the reader compares with "I_Only" and falls through to readElements_Q
    fi = api.file_probe(xml_bytes(body, "2048000", "Hz", "float32", 32, "LSB", channels=("I",)))
    assert fi.fmt.iq_order == "Q_Only" and fi.n_samples == len(body) // 4
    with pytest.raises(api.DabstarError):  # no data block: XmlDescriptor reports !ok
        api.file_probe(b"<?xml version='1.0'?><SDR><Sample><Samplerate Value='2048000'/></Sample></SDR>" + bytes(3000))


def test_probe_raw():
    fi = api.file_probe(np.full(20000, 127, np.uint8))
    assert (fi.kind, fi.sample_rate, fi.fmt.container, fi.fmt.iq_order, fi.data_offset, fi.n_samples) == ("raw", 2048000, "uint8", "IQ", 0, 10000)


# ------------------------------------------------------------------------------------------------ oracle restatements (CPU)
def numpy_resample(x, rate, reader):
    """out[2048 b + j] = x[b N + base_j + 1 - s] * r_j + x[b N + base_j - s] * (1 - r_j) in float32, s = 1 (XML) / 0 (WAV)."""
    n = rate // 1000
    i = np.arange(2048)
    if reader == "xml":
        in_val = np.float32(rate // 1000)
        base = np.floor(i * (np.float64(in_val) / 2048.0)).astype(np.int64)
        frac = (i.astype(np.float32) * (in_val / np.float32(2048.0)) - base.astype(np.float32)).astype(np.float32)
        blocks, shift = x.size // n, 1
    else:
        in_val = np.float32(rate) / np.float32(1000.0)
        base = np.floor(i.astype(np.float32) * (in_val / np.float32(2048.0))).astype(np.int64)
        frac = (i.astype(np.float32) * (in_val / np.float32(2048.0)) - base.astype(np.float32)).astype(np.float32)
        blocks, shift = (x.size - 1) // n if x.size else 0, 0
    xp = np.concatenate([np.zeros(1, np.complex64), x.astype(np.complex64)])  # xp[k + 1] = x[k], x[-1] = 0
    lo = (np.arange(blocks)[:, None] * n + base[None, :] - shift).reshape(-1)
    r = np.tile(frac, blocks)
    a, b = xp[lo + 2], xp[lo + 1]
    one_minus = (np.float32(1.0) - r).astype(np.float32)
    re = (a.real * r).astype(np.float32) + (b.real * one_minus).astype(np.float32)
    im = (a.imag * r).astype(np.float32) + (b.imag * one_minus).astype(np.float32)
    return (re.astype(np.float32) + 1j * im.astype(np.float32)).astype(np.complex64)


RATES = [2500000, 1792000, 2000000, 3000000, 1536000, 2048999]


@pytest.mark.parametrize("rate", RATES)
@pytest.mark.parametrize("reader", ["xml", "wav"])
def test_oracle_resample_matches_formula(oracle, rate, reader):
    rng = np.random.default_rng(rate % 1000 + len(reader))
    x = (rng.normal(size=20 * (rate // 1000) + 37) + 1j * rng.normal(size=20 * (rate // 1000) + 37)).astype(np.complex64)
    got, want = oracle.resample(x, rate, reader), numpy_resample(x, rate, reader)
    assert got.size == want.size == 20 * 2048
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("rate", RATES)
def test_xml_reader_resampling_against_the_reference(oracle, refo, rate):
    """The reference's own XmlReader thread (oracle/_ref) at a non-native sample rate: 1 ms blocks, linear interpolation
    (xml_reader.cpp:70-76, 212-231). Bit exact."""
    rng = np.random.default_rng(rate % 977)
    ms = 30
    n = ms * (rate // 1000)
    x16 = rng.integers(-20000, 20000, (n + 4 * (rate // 1000), 2)).astype(np.int16)
    got = refo.xml_reader_run(x16.view(np.uint8).reshape(-1), rate, "int16", 16, "LSB", "IQ", n)
    x = (x16.astype(np.float32) / np.float32(32768)).view(np.complex64).reshape(-1)
    want = oracle.resample(x, rate, "xml")
    m = ms * 2048
    assert got.size >= m and want.size >= m
    assert np.array_equal(got[:m].view(np.uint32), want[:m].view(np.uint32))


@pytest.mark.parametrize("rate", RATES)
def test_wav_reader_resampling_against_the_reference(oracle, refo, rate):
    """The reference's own WavReader thread over the float frames libsndfile would deliver (wav_reader.cpp:66-83, 196-211).
    Its interpolation table is computed in single precision under -ffast-math: gcc turns sampleRate / 1000.0f into a product
    with 0.001f, which moves the interpolation positions by up to 2e-4 of a sample unless the rate is exact in both forms; the
    restatement keeps the IEEE division. Same blocks, values within that; bit exact at 2.5 MS/s. The reader drops the last,
    incomplete 32768-frame read."""
    rng = np.random.default_rng(rate % 971)
    n_frames = 4 * 32768 + 1000
    x = ((rng.normal(size=n_frames) + 1j * rng.normal(size=n_frames)) * 0.3).astype(np.complex64)
    got = refo.wav_reader_run(x, rate)
    want = oracle.resample(x[:4 * 32768], rate, "wav")
    assert got.size == want.size == ((4 * 32768 - 1) // (rate // 1000)) * 2048
    assert np.abs(got - want).max() < 2e-3
    if rate == 2500000:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_oracle_pcm_conversions(oracle):
    rng = np.random.default_rng(5)
    u8 = rng.integers(0, 256, 2 * 500, dtype=np.uint8)
    want = ((u8.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)).view(np.complex64)
    assert np.array_equal(oracle.convert_samples(u8, 6, 0, 0, 0, 500), want)
    i32 = rng.integers(-2 ** 31, 2 ** 31, 2 * 500).astype(np.int32)
    want = (i32.astype(np.float32) / np.float32(2.0 ** 31)).view(np.complex64)
    assert np.array_equal(oracle.convert_samples(i32.view(np.uint8), 7, 0, 0, 0, 500), want)


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("rate", RATES)
@pytest.mark.parametrize("reader", ["xml", "wav"])
def test_resample_bit_exact(ctx, oracle, rate, reader):
    rng = np.random.default_rng(rate % 977)
    n = 53 * (rate // 1000) + 11
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    got = ctx.resample_linear(x, rate, api.READER_XML if reader == "xml" else api.READER_WAV)
    want = oracle.resample(x, rate, reader)
    assert got.size == want.size and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_resample_edges(ctx):
    x = np.arange(5000).astype(np.complex64)
    assert np.array_equal(ctx.resample_linear(x, 2048000), x)              # the readers pass 2.048 MS/s through
    assert ctx.resample_linear(x[:100], 2500000).size == 0                  # less than one 1 ms block
    assert ctx.resample_linear(x[:2500], 2500000, api.READER_XML).size == 2048
    assert ctx.resample_linear(x[:2500], 2500000, api.READER_WAV).size == 0  # the WAV reader's first block needs N + 1 samples
    with pytest.raises(api.DabstarError):
        ctx.resample_linear(x, 100)


@pytest.mark.gpu
def test_pcm_containers_bit_exact(ctx, oracle):
    rng = np.random.default_rng(8)
    u8 = rng.integers(0, 256, 2 * 3000, dtype=np.uint8)
    assert np.array_equal(ctx.ingest_convert(u8, api.SampleFormat("uint8_pcm")), oracle.convert_samples(u8, 6, 0, 0, 0, 3000))
    i32 = rng.integers(-2 ** 31, 2 ** 31, 2 * 3000).astype(np.int32)
    assert np.array_equal(ctx.ingest_convert(i32, api.SampleFormat("int32_pcm")), oracle.convert_samples(i32.view(np.uint8), 7, 0, 0, 0, 3000))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["wav", "xml"])
def test_resampled_recording_end_to_end(ctx, oracle, kind):
    """A 2.5 MS/s int16 recording in a WAV / UFF container: probe -> convert -> resample -> decode on the GPU equals the CPU
    chain on the oracle's conversion and resampling of the same file."""
    from scipy.signal import resample_poly
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(8, seed=77, snr_db=25.0, subch=sc, fmt=synth.FMT_CF32)
    hi = resample_poly(rec.iq.astype(np.complex128), 625, 512)              # 2.048 -> 2.5 MS/s
    pcm = np.clip(np.round(np.stack([hi.real, hi.imag], axis=1) * 8000.0), -32768, 32767).astype(np.int16)
    if kind == "wav":
        f = wav_bytes(pcm, 2500000, 1, 16)
    else:
        f = xml_bytes(pcm.tobytes(), "2500000", "Hz", "int16", 16, "LSB")
    x, fi = ctx.read_file(f)
    assert fi.kind == kind and fi.sample_rate == 2500000 and fi.n_samples == pcm.shape[0]
    want_x = oracle.resample(oracle.convert_samples(pcm.view(np.uint8), 2, 16, 0, 0, pcm.shape[0]), 2500000, kind)
    assert np.array_equal(x.view(np.uint32), want_x.view(np.uint32))
    want = oracle.chain_run(want_x, synth.subch_table(sc), 1)
    dp = api.DabProcessor(1, input_format=api.FMT_CF32, ctx=ctx)
    dp.set_audio_channel(0, sc)
    dp.run([x])
    got = dp.result(0)
    assert got.n_frames == want.n_frames >= 7
    assert np.array_equal(got.fic_valid, want.fic_valid) and got.fic_valid.mean() > 0.8
    ok = want.fic_valid.astype(bool).repeat(768, axis=1)
    assert np.array_equal(got.fib_bits[ok], want.fib_bits[ok])
    assert np.array_equal(got.msc[3], want.msc[3])
