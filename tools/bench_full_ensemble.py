#!/usr/bin/env python
"""configs[2] in numbers (not the bench line): full-ensemble decode, 18 mixed EEP / UEP sub-channels filling 864 CU, R recordings
of F frames resident in HBM, every sub-channel of every CIF through time de-interleave, depuncture, Viterbi and energy dispersal.
Prints one JSON line. Usage: python tools/bench_full_ensemble.py [--recordings 96] [--frames 104] [--unique 4]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SPEC = [(0, 0, 72, 108), (0, 1, 72, 72), (0, 2, 72, 54), (0, 3, 72, 36), (0, 4, 64, 54), (0, 5, 64, 42), (0, 6, 64, 36), (0, 7, 64, 30),
        (1, 3, 128, 96), (1, 4, 128, 84), (1, 5, 128, 64), (0, 2, 48, 36), (0, 2, 32, 24), (0, 6, 32, 18), (1, 5, 32, 16), (0, 3, 32, 16),
        (0, 2, 8, 6), (1, 4, 64, 42)]


def main():
    import torch
    from dabstar_b200 import api, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--recordings", type=int, default=96)
    ap.add_argument("--frames", type=int, default=104)
    ap.add_argument("--unique", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    subch, cu = [], 0
    for i, (sf, lvl, br, size) in enumerate(SPEC):
        subch.append(synth.SubChannel(i + 1, cu, size, sf, lvl, br))
        cu += size
    uniq = [synth.generate(args.frames, seed=900 + i, snr_db=15.0, subch=subch, fmt=synth.FMT_U8) for i in range(args.unique)]
    host = torch.from_numpy(np.stack([uniq[r % args.unique].iq for r in range(args.recordings)]))
    dev = host.cuda()
    stream = torch.cuda.Stream()
    ctx = api.Context(0, stream=stream)
    dp = api.DabProcessor(args.recordings, input_format=api.FMT_U8, max_window=128, ctx=ctx)
    for r in range(args.recordings):
        dp.set_audio_channel(r, subch)
    ptrs = [dev[r].data_ptr() for r in range(args.recordings)]
    ns = [host.shape[1]] * args.recordings
    with torch.cuda.stream(stream):
        dp.run_ptrs(ptrs, ns, api.MEM_DEVICE)
        res = dp.result(0)
        ok = all(np.array_equal(res.msc[s.sub_ch_id], uniq[0].msc_truth[j][:res.msc[s.sub_ch_id].shape[0]]) for j, s in enumerate(subch))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        acc = {}
        for _ in range(args.steps):
            dp.run_ptrs(ptrs, ns, api.MEM_DEVICE)
            for k, (ms, ln) in dp.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + ms
            gm, tm = dp.msc_kernel_ms()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    frames = sum(dp.result(r).n_frames for r in range(args.recordings))
    info_bits = frames * (3072 + 4 * 24 * sum(s.bit_rate for s in subch))
    print(json.dumps({"workload": "configs[2] full ensemble, 18 sub-channels (864 CU)", "recordings": args.recordings, "frames_per_step": frames,
                      "ms_per_step": ms, "frames_per_s": frames / ms * 1e3, "decoded_mbit_s": info_bits / ms / 1e3, "payload_equals_transmitted": bool(ok), "msc_gather_ms": gm, "msc_trellis_ms": tm,
                      "stages_ms": {k: v / args.steps for k, v in acc.items() if v > 0}}))


if __name__ == "__main__":
    main()
