"""The bundled transmitter against the oracle receiver: what goes in comes out (encode -> channel -> decode)."""
import numpy as np
import pytest

from dabstar_b200 import synth


@pytest.mark.parametrize("fmt", [synth.FMT_CF32, synth.FMT_U8, synth.FMT_I16])
def test_round_trip(oracle, fmt):
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(7, 400, 64, 1, 5, 128)]
    rec = synth.generate(7, seed=3, snr_db=25.0, subch=sc, fmt=fmt)
    iq = oracle.to_cf32(rec.iq)
    r = oracle.chain_run(iq, synth.subch_table(sc), 2)
    assert r.n_frames == 7 and r.fic_valid.all() and r.n_good_fibs == 84
    assert np.array_equal(r.fib_bits, rec.fib_truth)
    for i, s in enumerate(sc):
        assert r.msc[s.sub_ch_id].shape[0] == 4 * 7 - 16
        assert np.array_equal(r.msc[s.sub_ch_id], rec.msc_truth[i][:12])


def test_deterministic_and_seeded():
    a = synth.generate(2, seed=9, snr_db=10.0, fmt=synth.FMT_U8)
    b = synth.generate(2, seed=9, snr_db=10.0, fmt=synth.FMT_U8)
    c = synth.generate(2, seed=10, snr_db=10.0, fmt=synth.FMT_U8)
    assert np.array_equal(a.iq, b.iq) and not np.array_equal(a.iq, c.iq)
    assert a.iq.shape[0] == 60000 + 2 * 196608 + 4096


def test_rejects_bad_subchannel():
    with pytest.raises(ValueError):
        synth.generate(1, subch=[synth.SubChannel(1, 860, 54, 0, 2, 72)])
    with pytest.raises(ValueError):
        synth.generate(1, subch=[synth.SubChannel(1, 0, 54, 1, 1, 56)])  # no UEP profile 56 kbit/s level 1
