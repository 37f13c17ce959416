"""Row I1: file sample formats. CPU part: the oracle's restatement of XmlReader::readElements_* against independent numpy
formulas (xml_reader.cpp:254-800) and against the reference's OWN XmlReader thread (compiled into oracle/_ref with stand-ins for
its Qt parent widget and the XML descriptor parser): bit exact for every container / byte order / channel order except the
reader defects listed in test_xml_reader_against_the_reference, which are asserted as such. GPU part: dabstar_ingest_convert bit exact against the oracle, and a recording stored as
big-endian QI int16 decodes to the same bits as its native little-endian IQ form."""
import itertools

import numpy as np
import pytest

from dabstar_b200 import api, synth

WIDTH = {"int8": 1, "uint8": 1, "int16": 2, "int24": 3, "int32": 4, "float32": 4}
CASES = [(c, bo, od) for c, bo, od in itertools.product(WIDTH, ("LSB", "MSB"), api.IQ_ORDERS) if not (WIDTH[c] == 1 and bo == "MSB")]


def raw_bytes(container, n_values, seed):
    rng = np.random.default_rng(seed)
    if container == "float32":
        return rng.normal(size=n_values).astype(np.float32).view(np.uint8)
    return rng.integers(0, 256, n_values * WIDTH[container], dtype=np.uint8)


def numpy_reference(raw, container, bits, byte_order, order, n):
    """Independent statement of the conversion in float32 numpy arithmetic."""
    w = WIDTH[container]
    per = 2 if order in ("IQ", "QI") else 1
    b = raw[:n * per * w].reshape(-1, w).astype(np.int64)
    if byte_order == "MSB":
        b = b[:, ::-1]
    u = sum(b[:, i] << (8 * i) for i in range(w))
    if container == "uint8":
        v = (u.astype(np.float32) - np.float32(127.38)) / np.float32(128.0)
    elif container == "int8":
        s = np.where(u >= 128, u - 256, u)
        v = (s.astype(np.float32) / np.float32(127.0)) if order == "IQ" else (s.astype(np.float64) / 127.0).astype(np.float32)
    elif container == "float32":
        v = u.astype(np.uint32).view(np.float32)
    else:
        s = np.where(u >= (1 << (8 * w - 1)), u - (1 << (8 * w)), u)
        nb = bits or 8 * w
        scaler = np.float32(-2.0 ** 31 if nb == 32 else 2.0 ** (nb - 1))  # i32 shift() wraps at 32 bits (xml_reader.cpp:43-51)
        v = s.astype(np.float32) / scaler
    out = np.zeros((n, 2), np.float32)
    if order == "IQ":
        out[:] = v.reshape(n, 2)
    elif order == "QI":
        out[:] = v.reshape(n, 2)[:, ::-1]
    elif order == "I_Only":
        out[:, 0] = v
    else:
        out[:, 1] = v
    return out.view(np.complex64).reshape(-1)


# What the reference's reader does differently from its own other branches (all in xml_reader.cpp); the product follows the
# formulas of the other branches (include/dabstar_b200.h documents the list).
REFERENCE_DEFECTS = {
    ("uint8", "LSB", "QI"): "table indexed with the loop counter instead of the sample byte (:421)",
    ("int8", "LSB", "Q_Only"): "real part set to 127 (:690)",
    ("int24", "MSB", "IQ"): "one byte taken from offset 4 i + 4 instead of 6 i + 4 (:309)",
    ("int24", "MSB", "QI"): "the same offset, and the sign extension ORs 0x7F000000 instead of 0xFF000000 (:456-463)",
    ("float32", "LSB", "QI"): "I and Q are not swapped (:536-543)",
    ("float32", "MSB", "QI"): "I and Q are not swapped (:525-533)",
}


@pytest.mark.parametrize("container,byte_order,order", CASES)
def test_xml_reader_against_the_reference(oracle, refo, container, byte_order, order):
    n = 2048 * 4   # the reader delivers 2048 samples per millisecond at the native rate
    per = 2 if order in ("IQ", "QI") else 1
    for bits in {"int16": (16, 12), "int24": (24, 20), "int32": (32, 24)}.get(container, (0,)):
        raw = raw_bytes(container, per * (n + 4096), seed=3 + bits)
        if container == "float32":   # finite floats only (bit patterns of NaNs do not survive a swap in numpy)
            raw = np.random.default_rng(bits).normal(size=per * (n + 4096)).astype(np.float32).view(np.uint8)
        got = refo.xml_reader_run(raw, 2048000, container, bits or 8 * WIDTH[container], byte_order, order, n)[:n]
        want = oracle.convert_samples(raw, api.CONTAINERS[container], bits, byte_order == "MSB", api.IQ_ORDERS[order], n)
        assert got.size == n
        key = (container, byte_order, order)
        if key in REFERENCE_DEFECTS:
            assert not np.array_equal(got.view(np.uint32), want.view(np.uint32)), REFERENCE_DEFECTS[key]
            if container == "float32":   # exactly: the reference treats QI like IQ
                as_iq = oracle.convert_samples(raw, api.CONTAINERS[container], bits, byte_order == "MSB", api.IQ_ORDERS["IQ"], n)
                assert np.array_equal(got.view(np.uint32), as_iq.view(np.uint32))
            if key == ("int8", "LSB", "Q_Only"):
                assert np.all(got.real == 127.0) and np.allclose(got.imag, want.imag, rtol=2e-7)
        elif container == "int8":
            # v / 127: the reference is built with -ffast-math (CMakeLists.txt:76), gcc multiplies by the reciprocal: one ulp
            d = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
            assert d.max() <= 1
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (key, bits)


@pytest.mark.parametrize("container,byte_order,order", CASES)
def test_oracle_conversion_matches_formulas(oracle, container, byte_order, order):
    n = 4096
    bits = {"int16": 12, "int24": 24, "int32": 0}.get(container, 0)
    raw = raw_bytes(container, 2 * n, seed=len(container) + len(order))
    got = oracle.convert_samples(raw, api.CONTAINERS[container], bits, byte_order == "MSB", api.IQ_ORDERS[order], n)
    want = numpy_reference(raw, container, bits, byte_order, order, n)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_oracle_native_formats_agree_with_chain_input(oracle):
    # the two formats the chain tests feed (to_cf32) are the same conversions
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, (1000, 2), dtype=np.uint8)
    i16 = rng.integers(-32768, 32768, (1000, 2)).astype(np.int16)
    assert np.array_equal(oracle.convert_samples(u8, 1, 0, 0, 0, 1000), oracle.to_cf32(u8))
    assert np.array_equal(oracle.convert_samples(i16.view(np.uint8), 2, 16, 0, 0, 1000), oracle.to_cf32(i16))


@pytest.mark.gpu
@pytest.mark.parametrize("container,byte_order,order", CASES)
def test_ingest_convert_bit_exact(ctx, oracle, container, byte_order, order):
    n = 100_003  # not a multiple of the block size
    bits = {"int16": 14, "int24": 20, "int32": 32}.get(container, 0)
    raw = raw_bytes(container, 2 * n, seed=7 + len(container) + len(order))
    fmt = api.SampleFormat(container, bits, byte_order, order)
    got = ctx.ingest_convert(raw, fmt, n)
    want = oracle.convert_samples(raw, api.CONTAINERS[container], bits, byte_order == "MSB", api.IQ_ORDERS[order], n)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_ingest_rejects_unknown_formats(ctx):
    with pytest.raises(api.DabstarError):
        ctx.ingest_convert(np.zeros(16, np.uint8), api.SampleFormat("int12"))
    assert ctx.ingest_convert(np.zeros(0, np.uint8), api.SampleFormat("int16")).size == 0


@pytest.mark.gpu
def test_big_endian_qi_recording_decodes_like_native(ctx):
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(8, seed=21, snr_db=18.0, subch=sc, fmt=synth.FMT_I16)
    native = api.DabProcessor(1, input_format=api.FMT_I16, ctx=ctx)
    native.set_audio_channel(0, sc)
    native.run([rec.iq])
    a = native.result(0)
    # the same samples as an XML file would store them with Ordering=MSB, iqOrder=QI
    qi_be = np.ascontiguousarray(rec.iq[:, ::-1]).astype(">i2").view(np.uint8)
    conv = api.DabProcessor(1, input_format=api.FMT_CF32, ctx=ctx)
    conv.set_audio_channel(0, sc)
    conv.run_files([qi_be], api.SampleFormat("int16", 16, "MSB", "QI"))
    b = conv.result(0)
    assert a.n_frames == b.n_frames == 8
    assert np.array_equal(a.fib_bits, b.fib_bits) and np.array_equal(a.fic_valid, b.fic_valid)
    assert np.array_equal(a.msc[3], b.msc[3])
    assert [i.sym0_pos for i in a.info] == [i.sym0_pos for i in b.info]
