"""N>1 plumbing on CPU: world_size 2 over gloo. The data path has no collective; ranks only agree on the shard layout
and combine counts / times (dabstar_b200/parallel.py), which is what bench.py does under torchrun."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from dabstar_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rec, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_recordings(n_rec, rank, world)
    # pretend every recording decodes 10 + index frames and this rank needs (rank + 1) ms
    frames = sum(10 + r for r in mine)
    tot, ms = parallel.combine(frames, float(rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((tot, ms, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_all_recordings_once():
    world, n_rec = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_rec, q)) for r in range(world)]
    for p in ps:
        p.start()
    tot, ms, gathered = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for g in gathered for i in g)
    assert flat == list(range(n_rec))                    # every recording exactly once
    assert tot == sum(10 + r for r in range(n_rec))      # whole-job count
    assert ms == 2.0                                     # max over ranks


def test_frame_sharding_covers_the_stream():
    n = 37500
    seen = []
    for r in range(8):
        first, last, warm = parallel.shard_frames(n, r, 8)
        assert warm == (0 if r == 0 else 4)
        seen.extend(range(first, last))
    assert seen == list(range(n))
    assert parallel.shard_recordings(5, 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.combine(12, 3.5) == (12, 3.5)        # no process group: identity
