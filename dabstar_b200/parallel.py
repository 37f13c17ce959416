"""Multi-GPU plumbing: the path shards by recording with no data-path collective.
torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only to agree on the shard layout and to combine the
per-rank frame counts and device times."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_recordings(n_recordings: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of recordings to ranks (recordings are independent DabProcessors)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_recordings, world))


def shard_frames(n_frames: int, rank: int, world: int, warmup: int = 4) -> tuple[int, int, int]:
    """Contiguous frame range [first, last) of one long recording for this rank plus the number of warm-up frames to
    decode in front of it (>= 4 frames fill the 16-CIF time de-interleaver exactly; the demapper's IIRs need more and
    are approximate across a cut, SURVEY.md section 7 hard part 1). Returns (first, last, warmup_frames)."""
    per = (n_frames + world - 1) // world
    first = min(n_frames, rank * per)
    last = min(n_frames, first + per)
    return first, last, min(first, warmup)


def combine(frames: int, ms: float, device: torch.device | None = None) -> tuple[int, float]:
    """(sum of frames over ranks, max of time over ranks). Works without an initialised process group (world 1)."""
    if not (dist.is_available() and dist.is_initialized()):
        return frames, ms
    f = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(f.item()), float(t.item())
