"""N>1 path on CPU: view sharding is a pure partition (no collective on the render path); the only exchange is the
max-over-ranks of the timing and the gather of per-view checksums, exercised here with gloo, world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

import gel_b200


@pytest.mark.parametrize("n,world", [(8192, 8), (360, 7), (5, 8), (0, 4), (64, 1)])
def test_shards_partition_the_view_list(n, world):
    spans = [gel_b200.shard_views(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n = 37
    lo, hi = gel_b200.shard_views(n, world, rank)
    # stand-in for the per-view device checksums of this rank's block
    mine = np.arange(lo, hi, dtype=np.int64) * 7 + 3
    full = bench.gather_view_values(mine, n, world, rank)
    worst = bench.max_over_ranks(1.0 + rank)
    dist.barrier()
    q.put((rank, full.tolist() if full is not None else None, worst))
    dist.destroy_process_group()


def test_gloo_world2_gather_and_max():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1] == [k * 7 + 3 for k in range(37)]       # rank 0 holds every view's value, in view order
    assert res[0][2] == res[1][2] == 2.0                     # max over ranks


def _build_inputs_worker(args):
    workdir, name = args
    import bench
    d = bench.build_inputs(name, workdir)
    return int(d["tv"].shape[0]), int(d["tex"].shape[0])


def test_ranks_can_generate_the_bench_inputs_concurrently(tmp_path):
    """Under torchrun every rank calls bench.build_inputs on the same directory at once (a shared temporary name used to
    make one rank's rename fail): 8 processes, fresh directory, no leftovers."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(8) as pool:
        out = pool.map(_build_inputs_worker, [(str(tmp_path), "cfg1")] * 8)
    assert out == [(5000, 256)] * 8
    assert sorted(os.listdir(tmp_path)) == ["sphere50.obj", "tex256.bmp"]
