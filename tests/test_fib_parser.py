"""Row f1: multiplex configuration out of the FIC (FIG 0/0, 0/1, 0/2). The parser is host code in the product library, so
most of this runs without a GPU. The reference's FibDecoder needs Qt (QTimer, signals) and is not buildable here: this
boundary is pinned by EN 300 401 bit layouts as the reference's source states them (fib_decoder_fig0.cpp:89-290,
fib_table.h:51) and by round trips through the bundled transmitter's MCI generator."""
import numpy as np
import pytest

from dabstar_b200 import api, synth

ENSEMBLE = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(5, 200, 84, 1, 4, 128), synth.SubChannel(9, 400, 30, 0, 7, 64),
            synth.SubChannel(17, 0, 96, 1, 3, 128), synth.SubChannel(22, 500, 140, 1, 1, 128), synth.SubChannel(40, 700, 108, 0, 0, 72)]


def fib_bytes_to_bits(b: bytes) -> np.ndarray:
    assert len(b) <= 30
    raw = np.frombuffer(b + b"\xff" * (1 if len(b) < 30 else 0) + b"\x00" * max(0, 29 - len(b)), np.uint8)[:30]
    return np.concatenate([np.unpackbits(raw), np.zeros(16, np.uint8)])  # the CRC field is not looked at


def test_round_trip_through_the_transmitter():
    rec = synth.generate(3, seed=5, subch=ENSEMBLE, fig_mode=1, eid=0x4FFF, snr_db=100.0)
    p = api.FibParser()
    assert p.ensemble() is None and p.get_cif_count() == (-1, -1)
    p.process_FIB(rec.fib_truth[0].reshape(12, 256))
    got = p.sub_channels()
    assert [(s.sub_ch_id, s.start_cu, s.size_cu, s.short_form, s.prot_level, s.bit_rate) for s in got] == \
           [(s.sub_ch_id, s.start_cu, s.size_cu, s.short_form, s.prot_level, s.bit_rate) for s in ENSEMBLE]
    e = p.ensemble()
    assert (e.eid, e.cif_count_hi, e.cif_count_lo, e.change_flags) == (0x4FFF, 0, 0, 0)
    comps = p.components()
    assert [(c.sid, c.sub_ch_id, c.tmid, c.type, c.primary) for c in comps] == [(0xD000 + s.sub_ch_id, s.sub_ch_id, 0, 0 if s.short_form else 63, 1) for s in ENSEMBLE]
    # the following frames repeat the MCI: nothing is added twice, the CIF counter advances by 4 per frame
    p.process_FIB(rec.fib_truth[1].reshape(12, 256))
    p.process_FIB(rec.fib_truth[2].reshape(12, 256))
    assert len(p.sub_channels()) == len(ENSEMBLE) and len(p.components()) == len(ENSEMBLE)
    assert p.get_cif_count() == (0, 8) and p.ensemble().restarts == 0


def test_cif_counter_wraps_at_250():
    rec = synth.generate(70, seed=1, subch=ENSEMBLE[:1], fig_mode=1, snr_db=100.0, lead_samples=0, tail_samples=0)
    p = api.FibParser()
    for f in (62, 63, 69):
        p.process_FIB(rec.fib_truth[f].reshape(12, 256))
        assert p.get_cif_count() == divmod(4 * f, 250)


def test_long_and_short_form_bit_layout():
    # hand-assembled FIG 0/1 (EN 300 401 6.2.1): sub-channel 1 short form index 35 (128 kbit/s level 3, 96 CU) at CU 12,
    # sub-channel 2 long form option 1 level 1 (2-B), 42 CU at CU 300 -> 64 kbit/s
    fig = bytes([0x00 | 8, 0x01, (1 << 2) | 0, 12, 35, (2 << 2) | (300 >> 8), 300 & 0xff, 0x80 | (1 << 4) | (1 << 2) | 0, 42])
    p = api.FibParser()
    p.process_FIB(fib_bytes_to_bits(fig))
    a, b = p.sub_channels()
    assert (a.sub_ch_id, a.start_cu, a.size_cu, a.short_form, a.prot_level, a.bit_rate) == (1, 12, 96, 1, 3, 128)
    assert (b.sub_ch_id, b.start_cu, b.size_cu, b.short_form, b.prot_level, b.bit_rate) == (2, 300, 42, 0, 5, 64)


def test_table8_has_the_64_standard_rows():
    # every short-form index through the parser; spot values of EN 300 401 table 8
    want = {0: (16, 5, 32), 13: (52, 2, 56), 18: (70, 1, 64), 32: (104, 2, 112), 37: (140, 1, 128), 59: (208, 4, 320), 60: (280, 2, 320), 63: (416, 1, 384)}
    for idx, (cu, lvl, rate) in want.items():
        p = api.FibParser()
        p.process_FIB(fib_bytes_to_bits(bytes([4, 0x01, (7 << 2), 0, idx])))
        s, = p.sub_channels()
        assert (s.size_cu, s.prot_level, s.bit_rate) == (cu, lvl, rate), idx


def test_overlap_or_overflow_drops_the_database():
    ok = bytes([4, 0x01, (1 << 2), 10, 35])            # sub-channel 1: CU 10..105
    clash = bytes([4, 0x01, (2 << 2), 100, 35])        # sub-channel 2: CU 100..195 overlaps
    beyond = bytes([5, 0x01, (3 << 2) | 3, 0x20, 0x80 | (3 << 2), 100])  # CU 800 + 100 > 864
    p = api.FibParser()
    p.process_FIB(fib_bytes_to_bits(ok + clash))
    assert p.sub_channels() == [] and p.ensemble() is None
    p.process_FIB(fib_bytes_to_bits(ok))
    assert len(p.sub_channels()) == 1
    p.process_FIB(fib_bytes_to_bits(beyond))
    assert p.sub_channels() == []


def test_unknown_figs_are_skipped_and_garbage_is_harmless():
    fig1 = bytes([0x20 | 3, 0x00, 0x12, 0x34])          # FIG type 1, 3 bytes: skipped by its length
    fig013 = bytes([0x00 | 3, 13, 0xaa, 0xbb])          # FIG 0/13: skipped
    ok = bytes([4, 0x01, (1 << 2), 10, 35])
    p = api.FibParser()
    p.process_FIB(fib_bytes_to_bits(fig1 + fig013 + ok))
    assert len(p.sub_channels()) == 1
    rng = np.random.default_rng(0)
    p.process_FIB(rng.integers(0, 2, (500, 256), dtype=np.uint8))  # random FIBs: must not crash or loop
    p.process_FIB(fib_bytes_to_bits(bytes([0x1f]) + bytes(29)))     # length 31 in a 30-byte FIB


@pytest.mark.gpu
def test_self_configuration_decodes_like_a_configured_run(ctx, oracle):
    sc = ENSEMBLE[:4]
    rec = synth.generate(9, seed=12, snr_db=16.0, subch=sc, fmt=synth.FMT_U8, fig_mode=1, eid=0x10C4)
    dp = api.DabProcessor(1, input_format=api.FMT_U8, ctx=ctx)
    dp.set_auto_config(0)
    dp.start_eti_generator(0)
    dp.run([rec.iq])
    found = dp.sub_channels(0)
    assert [(s.sub_ch_id, s.start_cu, s.size_cu, s.short_form, s.prot_level, s.bit_rate, s.start_frame) for s in found] == \
           [(s.sub_ch_id, s.start_cu, s.size_cu, s.short_form, s.prot_level, s.bit_rate, 1) for s in sc]
    e = dp.ensemble(0)
    assert (e.eid, e.n_subch, e.cif_count_hi, e.cif_count_lo) == (0x10C4, 4, 0, 32)
    # the same recording through the CPU chain with the sub-channels given (Backends created in frame 1, CIF counter from FIG 0/0)
    given = [synth.SubChannel(s.sub_ch_id, s.start_cu, s.size_cu, s.short_form, s.prot_level, s.bit_rate, 1) for s in sc]
    want = oracle.chain_run(oracle.to_cf32(rec.iq), synth.subch_table(given), len(given), eti=True, eti_cif=(-1, 0))
    got = dp.result(0)
    assert got.n_frames == want.n_frames == 9 and got.fic_valid.all()
    for s in sc:
        assert np.array_equal(got.msc[s.sub_ch_id], want.msc[s.sub_ch_id]) and got.msc[s.sub_ch_id].shape[0] == 4 * 8 - 16
    eti = dp.eti(0)
    assert np.array_equal(eti, want.eti.reshape(-1, 6144))
    assert list(eti[:6, 4]) == [16, 17, 18, 19, 20, 21]  # FCT = CIF counter of FIG 0/0 (frame 4 starts at CIF 16) + CIF in frame
