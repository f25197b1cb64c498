"""CPU tests (gloo, world_size 2) of the multi-rank plumbing: slab arithmetic and the one exchange step."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_ranges_partition_the_blocks(pkg):
    from pdynamo_mirror_b200.parallel import block_range
    for nblocks in (1, 7, 21, 737, 34992):
        for nranks in (1, 2, 3, 4, 8):
            seen = []
            for r in range(nranks):
                lo, hi = block_range(nblocks, r, nranks)
                seen += list(range(lo, hi))
            assert seen == list(range(nblocks))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import pdynamo_mirror_b200  # noqa: F401
    from pdynamo_mirror_b200.parallel import reduce_results
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    e, dm, g = rng.random(6), rng.random((3, 3)), torch.from_numpy(rng.random((50, 3)))
    reduce_results(e, dm, g)
    q.put((rank, e, dm, g.numpy()))
    dist.destroy_process_group()


def test_reduce_results_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_e = sum(np.random.default_rng(100 + r).random(6) for r in range(2))
    rngs = [np.random.default_rng(100 + r) for r in range(2)]
    parts = [(r.random(6), r.random((3, 3)), r.random((50, 3))) for r in rngs]
    for rank, e, dm, g in got:
        assert np.allclose(e, parts[0][0] + parts[1][0]) and np.allclose(e, exp_e)
        assert np.allclose(dm, parts[0][1] + parts[1][1])
        assert np.allclose(g, parts[0][2] + parts[1][2])
