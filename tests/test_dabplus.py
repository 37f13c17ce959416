"""Next row f2: the DAB+ outer code. CPU part: the oracle's Reed-Solomon / Fire-code restatement against the reference's own
objects (oracle/_ref), and the super-frame processor on generated streams. GPU part: dabstar_dabplus_decode against the oracle,
record for record and byte for byte, incl. uncorrectable words, lost and regained synchronisation."""
import numpy as np
import pytest

import helpers
from dabstar_b200 import api


def make_stream(oracle, bit_rate, n_sf, seed, lead=0, damage=None):
    rng = np.random.default_rng(seed)
    frames, truth = [], []
    for i in range(n_sf):
        bits, aus = helpers.dabplus_superframe(bit_rate, rng, oracle.rs_encode, dac_rate=(i + seed) % 2, sbr_flag=(i // 2 + seed) % 2)
        frames.append(bits)
        truth.append(aus)
    stream = np.concatenate(frames)
    if lead:
        stream = np.concatenate([rng.integers(0, 2, (lead, 24 * bit_rate), dtype=np.uint8), stream])
    if damage is not None:
        damage(stream, rng)
    return stream, truth


def test_rs_and_firecode_against_the_reference_objects(oracle, refo):
    rng = np.random.default_rng(11)
    for t in range(1500):
        data = rng.integers(0, 256, 110, dtype=np.uint8)
        cw = oracle.rs_encode(data)
        assert np.array_equal(cw, refo.rs_encode(data))
        ne = int(rng.integers(0, 9))
        cw[rng.choice(120, ne, replace=False)] ^= rng.integers(1, 256, ne, dtype=np.uint8)
        if t % 9 == 0:
            cw = rng.integers(0, 256, 120, dtype=np.uint8)
        (ra, a), (rb, b) = oracle.rs_decode(cw), refo.rs_decode(cw)
        assert ra == rb and np.array_equal(a, b), (t, ne)
        if ne <= 5 and t % 9:
            assert np.array_equal(a, data) and ra <= ne
    for t in range(4000):
        x = rng.integers(0, 256, 11, dtype=np.uint8)
        if t % 2:  # a valid header with one burst of up to 6 bits
            par = helpers._firecode_parity(x[2:].tobytes())
            x[0], x[1] = par >> 8, par & 0xFF
            assert oracle.firecode_check(x)
            burst = int(rng.integers(1, 64)) << int(rng.integers(0, 82))
            x ^= np.frombuffer(burst.to_bytes(11, "big"), np.uint8)
        assert oracle.firecode_check(x) == refo.firecode_check(x)
        (oa, ya), (ob, yb) = oracle.firecode_check_and_correct(x), refo.firecode_check_and_correct(x)
        assert oa == ob and np.array_equal(ya, yb)


def expected_decoder_events(rec, pay):
    """What Mp4Processor hands to its AAC decoder for the super-frames the restatement reports (mp4processor.cpp:317-397):
    every CRC-good access unit of an accepted super-frame, a concealment call for every other one (concealment is on by
    default), nothing for a super-frame whose RS / Fire-code repair failed."""
    ev = []
    for r, p in zip(rec, pay):
        if not r.ok:
            continue
        params = r.dac_rate | (r.sbr_flag << 1) | (r.aac_channel_mode << 2) | (r.ps_flag << 3) | (r.mpeg_surround << 4)
        for u in range(r.num_aus):
            if r.au_state[u] == 1:
                ev.append(("au", params, p[r.au_start[u]:r.au_start[u + 1] - 2].tobytes()))
            else:
                ev.append(("conceal", 960 * (2 if r.sbr_flag else 1) * 2))
    return ev


@pytest.mark.parametrize("bit_rate,scenario", [(72, "clean"), (72, "byte_errors"), (32, "heavy"), (128, "resync"), (48, "too_short"), (8, "junk"), (96, "lost_sync")])
def test_superframe_state_machine_against_the_reference_mp4processor(oracle, refo, bit_rate, scenario):
    """Pins the restated super-frame processor (oracle/dab_outer.c) to the reference's own Mp4Processor object: synchronisation
    on the Fire code, the 4-super-frame grace counter, the one-frame slide after a miss, RS repair, AU borders and CRCs."""
    def byte_errors(stream, rng, rate):
        by = np.packbits(stream.reshape(-1))
        pos = rng.choice(by.size, max(1, int(by.size * rate)), replace=False)
        by[pos] ^= rng.integers(1, 256, pos.size, dtype=np.uint8)
        stream[:] = np.unpackbits(by).reshape(stream.shape)

    def resync(stream, rng):
        stream[12:22] = rng.integers(0, 2, stream[12:22].shape, dtype=np.uint8)
        byte_errors(stream[30:], rng, 0.01)

    def lost_sync(stream, rng):  # five super-frames of noise: the grace counter runs out and the processor searches again
        stream[10:36] = rng.integers(0, 2, stream[10:36].shape, dtype=np.uint8)

    damage = {"clean": None, "byte_errors": lambda s, r: byte_errors(s, r, 0.02), "heavy": lambda s, r: byte_errors(s, r, 0.06), "resync": resync,
              "too_short": None, "junk": None, "lost_sync": lost_sync}[scenario]
    n_sf = 1 if scenario == "too_short" else 12
    stream, _ = make_stream(oracle, bit_rate, n_sf, seed=len(scenario) + bit_rate, lead=0 if scenario == "too_short" else 2, damage=damage)
    if scenario == "too_short":
        stream = stream[:4]
    if scenario == "junk":
        stream = np.random.default_rng(5).integers(0, 2, stream.shape, dtype=np.uint8)
    if scenario == "resync":
        stream = np.concatenate([stream[:40], np.random.default_rng(1).integers(0, 2, (1, 24 * bit_rate), dtype=np.uint8), stream[40:]])
    rec, pay = oracle.dabplus_run(stream, bit_rate)
    got = refo.mp4_events(stream, bit_rate)
    want = expected_decoder_events(rec, pay)
    assert len(got) == len(want), (len(got), len(want))
    assert got == want
    if scenario in ("clean", "byte_errors", "resync", "lost_sync"):
        assert sum(1 for e in got if e[0] == "au") > 10


@pytest.mark.parametrize("bit_rate", [32, 72, 128])
def test_superframe_processor_recovers_the_access_units(oracle, bit_rate):
    def damage(stream, rng):  # up to 5 byte errors in every code word: all corrected
        by = np.packbits(stream.reshape(-1))
        pos = rng.choice(by.size, by.size // 40, replace=False)
        by[pos] ^= rng.integers(1, 256, pos.size, dtype=np.uint8)
        stream[:] = np.unpackbits(by).reshape(stream.shape)
    stream, truth = make_stream(oracle, bit_rate, 6, seed=bit_rate, lead=3)
    rec, pay = oracle.dabplus_run(stream, bit_rate)
    assert [r.first_frame for r in rec] == [3 + 5 * i for i in range(6)] and all(r.ok for r in rec)
    for r, p, aus in zip(rec, pay, truth):
        assert api.Mp4Processor.access_units(r, p) == aus and r.rs_corrections == 0 and r.rs_errors == 0
    noisy, truth = make_stream(oracle, bit_rate, 6, seed=bit_rate + 1, lead=0, damage=damage)
    rec, pay = oracle.dabplus_run(noisy, bit_rate)
    good = [r for r in rec if r.ok]
    assert len(good) >= 5 and sum(r.rs_corrections for r in rec) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("bit_rate,scenario", [(72, "clean"), (72, "byte_errors"), (32, "heavy"), (128, "resync"), (48, "too_short")])
def test_dabplus_decode_matches_the_oracle(ctx, oracle, bit_rate, scenario):
    def byte_errors(stream, rng, rate):
        by = np.packbits(stream.reshape(-1))
        pos = rng.choice(by.size, max(1, int(by.size * rate)), replace=False)
        by[pos] ^= rng.integers(1, 256, pos.size, dtype=np.uint8)
        stream[:] = np.unpackbits(by).reshape(stream.shape)

    def resync(stream, rng):  # two super-frames destroyed in the middle, then one extra junk frame shifts the alignment
        stream[12:22] = rng.integers(0, 2, stream[12:22].shape, dtype=np.uint8)
        byte_errors(stream[30:], rng, 0.01)

    damage = {"clean": None, "byte_errors": lambda s, r: byte_errors(s, r, 0.02), "heavy": lambda s, r: byte_errors(s, r, 0.06), "resync": resync,
              "too_short": None}[scenario]
    n_sf = 1 if scenario == "too_short" else 12
    stream, _ = make_stream(oracle, bit_rate, n_sf, seed=len(scenario) + bit_rate, lead=0 if scenario == "too_short" else 2, damage=damage)
    if scenario == "too_short":
        stream = stream[:4]
    if scenario == "resync":
        stream = np.concatenate([stream[:40], np.random.default_rng(1).integers(0, 2, (1, 24 * bit_rate), dtype=np.uint8), stream[40:]])
    want_rec, want_pay = oracle.dabplus_run(stream, bit_rate)
    got_rec, got_pay = api.Mp4Processor(bit_rate, ctx).add_frames(stream)
    assert len(got_rec) == len(want_rec)
    for g, w in zip(got_rec, want_rec):
        assert (g.first_frame, g.ok, g.rs_errors, g.rs_corrections, g.fc_corrected, g.dac_rate, g.sbr_flag, g.aac_channel_mode, g.ps_flag, g.mpeg_surround, g.num_aus,
                tuple(g.au_start), tuple(g.au_state)) == w.key()
    assert np.array_equal(got_pay, want_pay)
    if scenario == "clean":
        assert all(r.ok for r in got_rec) and len(got_rec) == 12
    if scenario == "byte_errors":  # acquisition needs an undamaged raw header (FirecodeChecker::check), so early super-frames may be passed over
        assert all(r.ok for r in got_rec) and 9 <= len(got_rec) <= 12 and sum(r.rs_corrections for r in got_rec) > 50
    if scenario == "heavy":
        assert any(r.rs_errors for r in got_rec)
    if scenario == "resync":
        assert any(not r.ok for r in got_rec) and got_rec[-1].ok and (got_rec[-1].first_frame - 2) % 5 == 1  # realigned after the extra frame
