#!/usr/bin/env python
"""Debug aid: one long recording (device-resident) through the segmented decode with DABSTAR_TRACE=1."""
import os, sys, time
os.environ["DABSTAR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dabstar_b200 import api, synth
F, TILES = 104, int(sys.argv[1]) if len(sys.argv) > 1 else 192
rec = synth.generate(F, seed=2, snr_db=15.0, fmt=synth.FMT_U8)
u = torch.from_numpy(rec.iq).cuda()
L = F * 196608
dev = torch.cat([u[:60000]] + [u[60000:60000 + L]] * TILES + [u[60000 + L:]]).contiguous()
stream = torch.cuda.Stream()
ctx = api.Context(0, stream=stream)
dp = api.DabProcessor(1, input_format=api.FMT_U8, scan_mode=True, max_window=9984, ctx=ctx)
dp.set_segmentation(104, 18)
with torch.cuda.stream(stream):
    for i in range(3):
        t0 = time.time()
        dp.run_ptrs([dev.data_ptr()], [dev.shape[0]], api.MEM_DEVICE)
        torch.cuda.synchronize()
        print(f"run {i}: {1e3 * (time.time() - t0):.2f} ms wall, frames {dp.n_frames(0)}", file=sys.stderr)
