#!/usr/bin/env python
"""Debug aid: runs one synthetic recording through the CUDA decoder and the CPU oracle and prints per-frame control
values and soft-bit differences. Usage: python tools/gpu_debug_chain.py [cfo_hz] [snr_db] [n_frames] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dabstar_b200 import api, synth  # noqa: E402
from oracle_api import Oracle  # noqa: E402

cfo = float(sys.argv[1]) if len(sys.argv) > 1 else 31000.0
snr = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 12
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 5
sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
rec = synth.generate(nf, seed=seed, snr_db=snr, cfo_hz=cfo, subch=sc, fmt=synth.FMT_U8, lead_samples=77777)
o = Oracle("dabo")
iq_f = o.to_cf32(rec.iq)
want = o.chain_run(iq_f, synth.subch_table(sc), 1, tap_soft=True)
dp = api.DabProcessor(1, input_format=synth.FMT_U8)
dp.set_audio_channel(0, sc)
dp.run([rec.iq])
got = dp.result(0)
print("frames", got.n_frames, want.n_frames, "good fibs", got.n_good_fibs, want.n_good_fibs, "counters", got.counters)
for i in range(max(got.n_frames, want.n_frames)):
    for tag, r in (("gpu", got), ("cpu", want)):
        if i < r.n_frames:
            a = r.info[i]
            print(tag, i, a.sym0_pos, a.start_index, "fbb %.3f %.3f %.3f fsync %.3f cp %.5f ce %.4f" % (a.fbb_sym0, a.fbb_data, a.fbb_null, a.fsync, a.phase_cp, a.clock_err),
                  "ratio", a.fic_ratio_before, a.fic_ratio_after, "valid", list(r.fic_valid[i]))
    if i < min(got.n_frames, want.n_frames) and got.info[i].sym0_pos == want.info[i].sym0_pos:
        d = np.abs(dp.soft_bits(0, i).astype(np.int32) - want.soft_bits(i).astype(np.int32))
        print("   soft diff per symbol: max", d.max(axis=1)[:8], "... frac>1", (d > 1).mean(axis=1)[:8].round(4), "overall", (d > 1).mean())
for s in sc:
    a, b = got.msc[s.sub_ch_id], want.msc[s.sub_ch_id]
    print("msc", s.sub_ch_id, a.shape, b.shape, "equal rows", (a[:min(len(a), len(b))] == b[:min(len(a), len(b))]).all(axis=1).sum())
