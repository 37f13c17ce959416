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


# ---- one long recording as sample ranges (parallel.stream_shard / decode_stream_shard / stitch) with a stand-in decoder
T_F = parallel.T_FRAME
P0 = 60000 + 504          # symbol 0 of frame 0 in the pretend stream
N_FRAMES = 200
N_SAMPLES = 60000 + N_FRAMES * T_F + 4096


class _Info:
    def __init__(self, pos):
        self.sym0_pos = pos


class _FakeResult:
    def __init__(self, frames, base):
        import numpy as np
        self.n_frames = len(frames)
        self.info = [_Info(P0 + f * T_F - base) for f in frames]
        self.fib_bits = np.array([[f % 251] * 8 for f in frames], np.uint8).reshape(-1, 8)
        self.fic_valid = np.ones((len(frames), 4), np.uint8)
        # a Backend created with the run's first frame emits from its 17th CIF on
        cifs = [4 * f + c for f in frames for c in range(4)][16:]
        self.msc = {3: np.array([[g % 253] * 4 for g in cifs], np.uint8).reshape(-1, 4)}


class _FakeDp:
    """Decodes the frames that lie completely inside its input; a cold start loses the first two."""
    def __init__(self, base):
        self.base, self.frames = base, []

    def run(self, recs):
        lo, hi = recs[0]
        first = 0 if lo == 0 else (lo + 2 * T_F - P0 + T_F - 1) // T_F
        self.frames = [f for f in range(first, N_FRAMES) if P0 + f * T_F + T_F <= hi]
        return 1.0

    def result(self, r):
        return _FakeResult(self.frames, self.base)


class _Sub:
    sub_ch_id, start_frame = 3, 0


def _shard_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = parallel.stream_shard(N_SAMPLES, rank, world)
    part = parallel.decode_stream_shard(_FakeDp(sh.in_lo), (sh.in_lo, sh.in_hi), sh, [_Sub()])
    parts = parallel.gather(part)
    tot, ms = parallel.combine(part["frames"], part["ms"])
    if rank == 0:
        out.put((parallel.stitch(parts), tot, [p["decoded_frames"] for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_decode_one_stream_without_exchange():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got, tot, decoded = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # what one sequential run gives
    seq = _FakeDp(0)
    seq.run([(0, N_SAMPLES)])
    want = seq.result(0)
    assert tot == got["frames"] == want.n_frames == N_FRAMES
    assert got["pos"] == [i.sym0_pos for i in want.info]
    assert (got["fib"] == want.fib_bits).all() and (got["msc"][3] == want.msc[3]).all()
    assert decoded[1] > N_FRAMES // 2 + 20          # the second rank decoded its warm-up frames as well ...
    assert sum(decoded) - N_FRAMES < 40             # ... and that is all the redundancy there is


def test_stream_shards_cover_the_stream():
    n = 37500 * T_F + 64096
    prev_hi = 0
    for r in range(8):
        sh = parallel.stream_shard(n, r, 8)
        assert sh.own_lo == prev_hi and sh.in_lo <= sh.own_lo and sh.in_hi >= min(n, sh.own_hi)
        assert (sh.own_lo - sh.in_lo) == (0 if r == 0 else 25 * T_F)
        prev_hi = sh.own_hi
    assert prev_hi == n
    assert parallel.owned_frames([10, 10 + T_F, 10 + 2 * T_F], parallel.StreamShard(1, 2, T_F, 3 * T_F, 0, 4 * T_F)) == (1, 3)
    assert parallel.owned_msc_rows(5, 9, 100) == slice(4, 20)
    assert parallel.shard_recordings(5, 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.combine(12, 3.5) == (12, 3.5)        # no process group: identity
