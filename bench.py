#!/usr/bin/env python
"""bench.py — Mode-I frames/s decoded (IQ -> Viterbi) on N B200s, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): FIC-only decode of a ~10 000-frame synthetic Mode-I batch per GPU
(PRS sync, ingest + FFT, DQPSK demap of all 75 data symbols, FIC Viterbi, FIB CRC): 96 recordings of 104 frames
(10 s each, like configs[0]) of u8 IQ at 15 dB SNR. A step = one pass of the whole path over the batch.
  value : whole-job frames/s, inputs resident in HBM, CUDA-event time, max over ranks
  e2e   : the same through the public API with pinned HOST buffers (H2D of the IQ and D2H of the FIB bits inside)
  roofline / stages : per kernel family, algorithmic bytes over the CUDA-event time of its launches
  cpu_baseline : the reference's own CPU chain (oracle/_ref) or its C restatement on a bounded sample, rank 0, N=1
Scaling is weak: every rank decodes its own batch, there is no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_F = 196608
METRIC = "Mode I DAB frames/s decoded (IQ->Viterbi)"
UNIT = "frames/s"
# algorithmic bytes per frame of each kernel family (DESIGN.md section "Kernels")
BYTES_PER_FRAME = {
    "ingest_fft": 77 * 2048 * 2 + 77 * 1536 * 8,   # u8 IQ of the 77 useful parts in, nominal-carrier spectra out
    "demap": 77 * 1536 * 8 + 75 * 3072 * 2,        # spectra in, int16 soft bits out
    "cp_corr": 75 * 2 * 504 * 2,                   # both ends of every cyclic prefix
    "prs_corr": 2048 * 2,
}
ACS_PER_FRAME_FIC = 4 * 774 * 64


def env_int(name, default):
    return int(os.environ.get(name, default))


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_worker(args):
    """One process = one recording through the CPU chain (the reference keeps file-scope state: one process each)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dabstar_b200 import synth
    from oracle_api import Oracle
    o = Oracle(args.cpu_lib)
    rec = synth.generate(args.frames, seed=args.seed, snr_db=args.snr, fmt=synth.FMT_U8)
    iq = o.to_cf32(rec.iq)  # file reader conversion is not part of the timed chain (the reference does it in another thread)
    times, frames = [], 0
    for it in range(args.warmup + args.steps):
        r = o.chain_run(iq, scan_mode=1)
        if it >= args.warmup:
            times.append(r.seconds)
            frames += r.n_frames
        r.close()
    print(json.dumps({"frames": frames, "seconds": sum(times)}))


def run_cpu_chain(kind_lib: str, procs: int, frames: int, steps: int, warmup: int, snr: float):
    cmds = [[sys.executable, os.path.abspath(__file__), "--cpu-worker", "--cpu-lib", kind_lib, "--frames", str(frames), "--seed", str(1000 + i),
             "--steps", str(steps), "--warmup", str(warmup), "--snr", str(snr)] for i in range(procs)]
    ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for c in cmds]
    tot_frames, max_s = 0, 0.0
    for p in ps:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed: " + err[-400:])
        d = json.loads(out.strip().splitlines()[-1])
        tot_frames += d["frames"]
        max_s = max(max_s, d["seconds"])
    return tot_frames, max_s


def cpu_kind():
    from dabstar_b200 import build
    if os.path.exists(build.LIB_REF):
        return "reference", "dabref"
    build.build_oracle()
    return "port", "dabo"


def reference_arm(args, rank, world):
    if rank != 0:
        return
    kind, lib = cpu_kind()
    cores = max(1, min(os.cpu_count() or 1, 64))
    frames = 160
    t0 = time.time()
    tot, sec = run_cpu_chain(lib, cores, frames, args.steps, args.warmup, args.snr)
    value = tot / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": "configs[1] FIC-only decode, bounded CPU sample", "recordings": cores, "frames_per_recording": frames, "input": "u8 IQ 2.048 MS/s", "snr_db": args.snr},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{cores} processes x {frames} frames x {args.steps} steps, scalar Viterbi + scalar OfdmDecoder, FFT shim instead of FFTW ({time.time() - t0:.0f} s wall)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- native arm
# (short_form, prot_level, bit_rate, size_cu)
SWEEP_PROFILES = [("EEP 1-A 72k", 0, 0, 72, 108), ("EEP 2-A 72k", 0, 1, 72, 72), ("EEP 3-A 72k", 0, 2, 72, 54), ("EEP 4-A 72k", 0, 3, 72, 36),
                  ("EEP 1-B 64k", 0, 4, 64, 54), ("EEP 2-B 64k", 0, 5, 64, 42), ("EEP 3-B 64k", 0, 6, 64, 36), ("EEP 4-B 64k", 0, 7, 64, 30),
                  ("UEP 1 128k", 1, 1, 128, 140), ("UEP 2 128k", 1, 2, 128, 116), ("UEP 3 128k", 1, 3, 128, 96), ("UEP 4 128k", 1, 4, 128, 84),
                  ("UEP 5 128k", 1, 5, 128, 64)]


def viterbi_sweep(ctx, stream, n_frames, sm_mhz=1965.0):
    """Protection::deconvolve over n_frames logical frames per protection level; times 3 launches after 1 warm-up with CUDA events."""
    import ctypes
    import torch
    from dabstar_b200 import api
    out = {"logical_frames_per_level": n_frames, "levels": {}, "input": "noisy soft bits (+-60 with sigma 40), resident in HBM"}
    g = torch.Generator(device="cuda").manual_seed(4)
    tot_bits, tot_ms = 0.0, 0.0
    for name, sf, lvl, br, cu in SWEEP_PROFILES:
        n_soft = cu * 64
        soft = ((torch.randint(0, 2, (n_frames, n_soft), generator=g, device="cuda", dtype=torch.int16) * 2 - 1) * 60
                + (torch.randn((n_frames, n_soft), generator=g, device="cuda") * 40).to(torch.int16)).contiguous()
        bits = torch.empty((n_frames, 24 * br), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()

        def run():
            ctx.check(ctx.lib.dabstar_protection_deconvolve(ctx.h, sf, br, lvl, cu, ctypes.c_void_p(soft.data_ptr()), n_frames,
                                                            ctypes.c_void_p(bits.data_ptr()), api.MEM_DEVICE), "dabstar_protection_deconvolve")
        run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        info_bits = n_frames * 24 * br
        acs = n_frames * 64 * (24 * br + 6)
        # integer-ALU issue roofline: 148 SMs x 128 lanes x f_clk / 4 lane-ops per add-compare-select
        out["levels"][name] = {"ms": ms, "mbit_s": info_bits / ms / 1e3, "gacs": acs / ms / 1e6,
                               "frac_int_alu": (acs / ms / 1e6) / (148 * 128 * sm_mhz * 1e6 / 4.0 / 1e9)}
        tot_bits += info_bits
        tot_ms += ms
    out["mbit_s_overall"] = tot_bits / tot_ms / 1e3
    return out


def pin_to_gpu_cpus(local_rank, world):
    """One host control loop per GPU: keep each rank on the CPUs NVML reports as local to its GPU (its NUMA node), and inside
    that set on its own share, so that the ranks do not migrate across sockets or onto each other."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = [c for c in range(n_cpu) if (words[c // 64] >> (c % 64)) & 1]
        allowed = sorted(set(local) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        # ranks whose GPUs share this CPU set split it evenly
        same = [r for r in range(world) if list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r), (n_cpu + 63) // 64)) == list(words)]
        k, i = max(1, len(allowed) // max(1, len(same))), same.index(local_rank) if local_rank in same else 0
        mine = allowed[i * k:(i + 1) * k] or allowed
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


def native_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from dabstar_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pinned_cpus = pin_to_gpu_cpus(local_rank, world) if world > 1 and not args.no_cpu_affinity else None
    R, F = args.recordings, args.frames
    n_samples = 60000 + F * T_F + 4096

    # ---- synthetic recordings in pinned host memory; unique ones up to a time budget, then reused
    host = torch.empty((R, n_samples, 2), dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    t0 = time.time()
    unique = 0
    for r in range(R):
        if time.time() - t0 < args.synth_budget or unique == 0:
            synth.generate(F, seed=args.seed + 7919 * rank + r, snr_db=args.snr, fmt=synth.FMT_U8, out=hnp[r])
            unique += 1
        else:
            hnp[r] = hnp[r % unique]
    dev = host.cuda(non_blocking=False)
    stream = torch.cuda.Stream()
    ctx = api.Context(local_rank, stream=stream)
    dp = api.DabProcessor(R, input_format=api.FMT_U8, scan_mode=True, max_window=args.window, ctx=ctx)
    if args.segment_frames > 0:
        dp.set_segmentation(args.segment_frames, args.segment_warmup)
    d_ptrs = [dev[r].data_ptr() for r in range(R)]
    h_ptrs = [host[r].data_ptr() for r in range(R)]
    ns = [n_samples] * R

    def frames_done():
        return sum(dp.result(r).n_frames for r in range(R))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            dp.run_ptrs(d_ptrs, ns, api.MEM_DEVICE)
        frames_per_step = frames_done()
        cnt = np.sum([dp.result(r).counters for r in range(R)], axis=0)
        good_fibs = int(cnt[0])
        # ---- timed: inputs resident in HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks = ClockSampler(local_rank)
        barrier()
        clocks.start()
        launches0 = ctx.kernel_launches
        stage_acc = {}
        heavy_acc = [0.0, 0.0]
        e0.record(stream)
        for _ in range(args.steps):
            dp.run_ptrs(d_ptrs, ns, api.MEM_DEVICE)
            heavy_acc[0] += dp.heavy_ms(False)
            heavy_acc[1] += dp.heavy_ms(True)
            for k, (ms, ln) in dp.stage_ms().items():
                a = stage_acc.setdefault(k, [0.0, 0])
                a[0] += ms
                a[1] += ln
        e1.record(stream)
        barrier()
        clk = clocks.stop()
        ms_total = e0.elapsed_time(e1)
        launches = ctx.kernel_launches - launches0
        # ---- timed: end to end from pinned host memory (H2D of the IQ and D2H of the FIB bits inside the call)
        dp.run_ptrs(h_ptrs, ns, api.MEM_HOST)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            dp.run_ptrs(h_ptrs, ns, api.MEM_HOST)
            _ = dp.result(0).fib_bits
        e1.record(stream)
        barrier()
        ms_e2e = e0.elapsed_time(e1)

        # ---- configs[3]: Viterbi-only throughput per protection level (depuncture + K=7 decode of MSC logical frames, soft bits
        #      resident in HBM), the "Viterbi Mbit/s" half of the metric; bounded batch, rank 0's GPU only
        vit_sweep = None
        if rank == 0 and not args.no_viterbi_sweep:
            vit_sweep = viterbi_sweep(ctx, stream, args.viterbi_frames, clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    fr = torch.tensor([frames_per_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    ms_total, ms_e2e = t.tolist()
    total_frames = fr.item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = total_frames * args.steps / (ms_total / 1e3)
    e2e_value = total_frames * args.steps / (ms_e2e / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    stages = {}
    for k, (ms, ln) in stage_acc.items():
        if ln == 0:
            continue
        d = {"ms_per_step": ms / args.steps, "launches_per_step": ln / args.steps}
        if k in BYTES_PER_FRAME:
            gbs = BYTES_PER_FRAME[k] * frames_per_step * args.steps / (ms / 1e3) / 1e9
            d.update({"achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak})
        if k == "fic_viterbi":
            acs = ACS_PER_FRAME_FIC * frames_per_step * args.steps / (ms / 1e3)
            sm_mhz = clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0
            peak_acs = 148 * 128 * sm_mhz * 1e6 / 4.0   # 4 lane-ops per add-compare-select
            d.update({"achieved_gacs": acs / 1e9, "frac_int_alu": acs / peak_acs, "mbit_s": 3072 * frames_per_step * args.steps / (ms / 1e3) / 1e6})
        stages[k] = d
    # the FFT + demap STAGE as one span (its chunks overlap on two streams), against SURVEY.md 8(d)'s stage bytes: 76 symbols +
    # null symbol of spectra in, soft bits out = 1 722 368 B per frame
    stage_bytes = 77 * 2048 * 8 + 75 * 3072 * 2
    gbs = stage_bytes * frames_per_step * args.steps / (max(heavy_acc[0], 1e-9) / 1e3) / 1e9
    stages["fft_demap_stage"] = {"ms_per_step": heavy_acc[0] / args.steps, "with_fic_ms_per_step": heavy_acc[1] / args.steps, "algorithmic_bytes_per_frame": stage_bytes,
                                 "achieved_gbs_stage": gbs, "frac_hbm_stage": gbs / hbm_peak,
                                 "note": "first FFT launch to last demap launch of every window, CUDA events; the per-kernel times above overlap"}
    hbm_stages = {k: v for k, v in stages.items() if "achieved_gbs" in v}
    dom = max(hbm_stages, key=lambda k: hbm_stages[k]["ms_per_step"])
    roofline = {"kernel": {"ingest_fft": "k_fft_frames", "demap": "k_demap", "cp_corr": "k_cp_corr", "prs_corr": "k_prs_corr"}[dom], "bound": "hbm",
                "achieved": hbm_stages[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm_stages[dom]["frac_hbm"], "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_frame": BYTES_PER_FRAME[dom]}
    # achieved / traffic are per launch: algorithmic bytes of one launch over its average CUDA-event duration
    n_launch = max(1.0, hbm_stages[dom]["launches_per_step"])
    roofline["algorithmic_bytes_per_launch"] = BYTES_PER_FRAME[dom] * frames_per_step / n_launch
    roofline["ms_per_launch"] = hbm_stages[dom]["ms_per_step"] / n_launch
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        k = tr["kernels"].get(roofline["kernel"])
        if k and n_launch == 1.0 and int(tr["frames_per_launch"]) == int(frames_per_step):
            roofline["traffic"] = k["dram_read_bytes"] + k["dram_write_bytes"]
            roofline["traffic_source"] = tr["source"] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, same workload)"
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        kind, lib = cpu_kind()
        cores = max(1, min(os.cpu_count() or 1, 16))
        t0 = time.time()
        tot, sec = run_cpu_chain(lib, cores, 150, 2, 1, args.snr)
        cpu = {"value": tot / sec, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cores} processes x 150 frames x 2 passes of the same FIC-only workload ({time.time() - t0:.0f} s wall), scalar build, FFT shim instead of FFTW"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": f"synthetic ({unique} unique recordings per GPU of {R}, random FIB payloads, AWGN {args.snr} dB)",
        "config": {"workload": "configs[1] FIC-only decode of a 10k-frame batch", "recordings_per_gpu": R, "frames_per_recording": F,
                   "frames_per_step_all_gpus": total_frames, "input": "u8 IQ 2.048 MS/s", "input_bytes_per_gpu": int(R * n_samples * 2),
                   "l2": "inputs (3.9 GB) and intermediates far exceed the 126 MB L2", "window": args.window, "cpus_per_rank": pinned_cpus,
                   "fib_crc_pass": good_fibs / max(1.0, 12.0 * frames_per_step),
                   "windows_per_recording": float(cnt[4]) / R, "frames_through_heavy_pass": int(cnt[7]), "x_real_time": value / world / (2048000 / T_F)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(R * n_samples * 2), "d2h_bytes_per_step": int(frames_per_step * 384),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "stages": stages, "viterbi_sweep": vit_sweep, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--recordings", type=int, default=96)
    ap.add_argument("--frames", type=int, default=104)
    ap.add_argument("--window", type=int, default=128)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--snr", type=float, default=15.0)
    ap.add_argument("--synth-budget", type=float, default=45.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-viterbi-sweep", action="store_true")
    ap.add_argument("--no-cpu-affinity", action="store_true", help="N > 1: do not pin each rank to the CPUs local to its GPU")
    ap.add_argument("--viterbi-frames", type=int, default=131072, help="logical frames per protection level in the Viterbi-only sweep")
    ap.add_argument("--segment-frames", type=int, default=0, help="demap a recording's window as parallel segments of this many frames (0 = off)")
    ap.add_argument("--segment-warmup", type=int, default=18)
    ap.add_argument("--cpu-worker", action="store_true")
    ap.add_argument("--cpu-lib", default="dabo")
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker(args)
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun as the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    native_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
