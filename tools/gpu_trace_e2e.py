#!/usr/bin/env python
"""Debug aid: one host-input run of the bench workload with DABSTAR_TRACE=1 (per-round progress on stderr)."""
import os, sys, time
os.environ["DABSTAR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dabstar_b200 import api, synth
R, F = int(sys.argv[1]) if len(sys.argv) > 1 else 96, 104
n = 60000 + F * 196608 + 4096
host = torch.empty((R, n, 2), dtype=torch.uint8, pin_memory=True)
for r in range(min(R, 4)):
    synth.generate(F, seed=2 + r, snr_db=15.0, fmt=synth.FMT_U8, out=host.numpy()[r])
for r in range(4, R):
    host[r] = host[r % 4]
stream = torch.cuda.Stream()
ctx = api.Context(0, stream=stream)
dp = api.DabProcessor(R, input_format=api.FMT_U8, scan_mode=True, max_window=128, ctx=ctx)
ptrs = [host[r].data_ptr() for r in range(R)]
DEV_ONLY = len(sys.argv) > 2 and sys.argv[2] == "dev"
with torch.cuda.stream(stream):
    for i in range(0 if DEV_ONLY else 3):
        t0 = time.time()
        dp.run_ptrs(ptrs, [n] * R, api.MEM_HOST)
        torch.cuda.synchronize()
        print(f"run {i}: {1e3 * (time.time() - t0):.2f} ms wall, frames {sum(dp.result(r).n_frames for r in range(R))}", file=sys.stderr)
    dev_in = host.cuda()
    torch.cuda.synchronize()
    dptrs = [dev_in[r].data_ptr() for r in range(R)]
    for i in range(3):
        t0 = time.time()
        dp.run_ptrs(dptrs, [n] * R, api.MEM_DEVICE)
        torch.cuda.synchronize()
        print(f"device-resident run {i}: {1e3 * (time.time() - t0):.2f} ms wall", file=sys.stderr)
    del dev_in
    if DEV_ONLY:
        sys.exit(0)
    # plain copy speed for comparison
    dev = torch.empty_like(host, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    print(f"plain H2D of {host.numel() / 1e9:.2f} GB: {1e3 * (time.time() - t0):.2f} ms", file=sys.stderr)
