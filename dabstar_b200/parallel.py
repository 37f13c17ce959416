"""Multi-GPU plumbing. The path shards with NO data-path collective:

* independent recordings go round-robin over the ranks (`shard_recordings`);
* ONE long recording goes as contiguous sample ranges (`stream_shard`): a rank other than the first starts cold (time sync,
  coarse AFC) a few frames in front of its range, lets the demapper's per-carrier IIRs (ofdm_decoder.cpp:182-251) and
  Backend's 16-CIF de-interleaver (backend.cpp:129-161) warm up on those frames and throws their output away. A frame
  belongs to the rank whose range holds the position of its symbol 0, so every frame has exactly one owner without the
  ranks talking to each other; the decoded FIB / MSC bytes equal those of one sequential run wherever the PRS peak is
  unambiguous (tests/test_long_recordings.py). What differs is the soft bits: a cold start re-seeds the integer-Hz
  derotation (DabProcessor's f_sync accumulates a fraction of a hertz from the start of the stream, dab_processor.cpp:236-251).

torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for barriers, the max / sum of timing scalars and for
gathering results after the timed region."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

T_FRAME = 196608
T_U = 2048


def shard_recordings(n_recordings: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of recordings to ranks (recordings are independent DabProcessors)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_recordings, world))


@dataclass
class StreamShard:
    rank: int
    world: int
    own_lo: int    # frames whose symbol 0 starts in [own_lo, own_hi) are this rank's
    own_hi: int
    in_lo: int     # samples [in_lo, in_hi) are what this rank decodes
    in_hi: int

    @property
    def cold_start(self) -> bool:
        return self.in_lo > 0


def stream_shard(n_samples: int, rank: int, world: int, warmup_frames: int = 18, sync_frames: int = 3) -> StreamShard:
    """Sample range of one long recording for `rank`. In front of its own range a rank decodes
    sync_frames (level estimate of 20 x T_u samples + null-symbol search + first PRS / coarse-AFC frame)
    + warmup_frames (demapper IIRs) + 4 (16 CIFs of de-interleaver history) frames; behind it one more frame so that the
    last frame that starts inside the range is complete."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (n_samples * rank) // world
    hi = (n_samples * (rank + 1)) // world
    lead = (sync_frames + warmup_frames + 4) * T_FRAME
    return StreamShard(rank, world, lo, hi if rank + 1 < world else n_samples, max(0, lo - lead) if rank > 0 else 0,
                       min(n_samples, hi + T_FRAME + 2 * T_U) if rank + 1 < world else n_samples)


def owned_frames(sym0_positions, shard: StreamShard) -> tuple[int, int]:
    """[first, last) of this rank's decoded frames (positions relative to in_lo, as the decoder reports them) it owns."""
    pos = np.asarray(sym0_positions, np.int64) + shard.in_lo
    mine = np.nonzero((pos >= shard.own_lo) & (pos < shard.own_hi))[0]
    if mine.size == 0:
        return 0, 0
    first, last = int(mine[0]), int(mine[-1]) + 1
    if last - first != mine.size:
        raise RuntimeError("owned frames are not contiguous")
    return first, last


def owned_msc_rows(first: int, last: int, n_rows: int, backend_start_frame: int = 0) -> slice:
    """Rows of a sub-channel's payload (one per CIF, the Backend's first row is its 17th CIF, backend.cpp:146-150) that belong
    to frames [first, last) of the run."""
    row0 = 4 * backend_start_frame + 16
    a = max(0, 4 * first - row0)
    b = max(a, min(n_rows, 4 * last - row0))
    return slice(a, b)


def decode_stream_shard(dp, iq, shard: StreamShard, subch, device_ptr: int | None = None) -> dict:
    """Runs `dp` (an api.DabProcessor with one recording and `subch` set) over this rank's sample range and keeps what the rank
    owns. iq: the rank's samples [in_lo, in_hi) as a host array, or None with device_ptr = their device address."""
    n = shard.in_hi - shard.in_lo
    if device_ptr is not None:
        ms = dp.run_ptrs([device_ptr], [n], 1)
    else:
        ms = dp.run([iq])
    res = dp.result(0)
    first, last = owned_frames([i.sym0_pos for i in res.info], shard)
    if shard.cold_start and res.n_frames and first < 4:
        raise RuntimeError(f"rank {shard.rank}: only {first} frames in front of the owned range (the de-interleaver needs 4)")
    out = {"rank": shard.rank, "ms": ms, "frames": last - first, "decoded_frames": res.n_frames,
           "pos": [int(i.sym0_pos) + shard.in_lo for i in res.info[first:last]],
           "fib": res.fib_bits[first:last], "valid": res.fic_valid[first:last], "msc": {}}
    for s in subch:
        rows = res.msc[s.sub_ch_id]
        out["msc"][s.sub_ch_id] = rows[owned_msc_rows(first, last, rows.shape[0], s.start_frame if not shard.cold_start else 0)]
    return out


def stitch(parts: list[dict]) -> dict:
    """Concatenates the ranks' results in stream order."""
    parts = sorted(parts, key=lambda p: p["rank"])
    keys = parts[0]["msc"].keys()
    return {"frames": sum(p["frames"] for p in parts), "pos": [x for p in parts for x in p["pos"]],
            "fib": np.concatenate([p["fib"] for p in parts]), "valid": np.concatenate([p["valid"] for p in parts]),
            "msc": {k: np.concatenate([p["msc"][k] for p in parts]) for k in keys}}


def gather(part: dict) -> list[dict]:
    """All ranks' parts on every rank (after the timed region; object collective)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [part]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, part)
    return out


def combine(frames: int, ms: float, device: torch.device | None = None) -> tuple[int, float]:
    """(sum of frames over ranks, max of time over ranks). Works without an initialised process group (world 1)."""
    if not (dist.is_available() and dist.is_initialized()):
        return frames, ms
    f = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(f.item()), float(t.item())
