"""CPU tests (gloo, world_size 2) of the multi-rank plumbing: slab arithmetic, the result reduction and the two halo exchanges."""
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


def test_slab_ranges_tile_the_sorted_order(pkg):
    from pdynamo_mirror_b200.parallel import slab_range
    for n in (31, 648, 23558, 1119744):
        nblocks = (n + 31) // 32
        for nranks in (1, 2, 3, 8):
            edges = [slab_range(nblocks, n, r, nranks) for r in range(nranks)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(nranks - 1))


def _halo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import pdynamo_mirror_b200  # noqa: F401
    from pdynamo_mirror_b200.parallel import SlabExchange
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, slabs = 100, [(0, 48), (48, 100)]
    # rank 0 lists atoms 48..69 of slab 1; rank 1 lists atoms 0..9 of slab 0 (periodic wrap)
    mine = np.zeros((world, 2, 2), np.int64)               # [slab][lower / upper half] = [lo, hi)
    if rank == 0:
        mine[1, 0] = (48, 70)
    else:
        mine[0, 0] = (0, 10); mine[0, 1] = (40, 48)          # rank 1 reaches both ends of slab 0 (periodic wrap)
    ex = SlabExchange(rank, world)
    ex.set_ranges(mine, torch.device("cpu"))
    rng = np.random.default_rng(7 + rank)
    gs = torch.from_numpy(rng.random((n, 3)))
    ex.halo_to_owners(gs)
    xs = torch.full((n, 3), -1.0, dtype=torch.float64)
    s0, s1 = slabs[rank]
    xs[s0:s1] = torch.arange(s0, s1, dtype=torch.float64)[:, None] + 0.5 * rank
    halo = ex.owners_to_halo(xs)
    xa = xs.clone()
    ex.allgather_slabs(xa, slabs)
    q.put((rank, ex.table.copy(), gs.numpy(), xs.numpy(), halo, xa.numpy()))
    dist.destroy_process_group()


def test_halo_exchanges_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    part = [np.random.default_rng(7 + r).random((100, 3)) for r in range(2)]
    for rank in range(2):
        table, gs, xs, halo, xa = got[rank]
        assert table[0, 1, 0].tolist() == [48, 70] and table[1, 0, 0].tolist() == [0, 10] and table[1, 0, 1].tolist() == [40, 48]
        assert not table[0, 0].any() and not table[1, 1].any()
    # gradients: the owner's slab holds its own partial plus the other rank's halo contribution, nothing else moved
    g0, g1 = got[0][1], got[1][1]
    exp0 = part[0].copy(); exp0[0:10] += part[1][0:10]; exp0[40:48] += part[1][40:48]
    exp1 = part[1].copy(); exp1[48:70] += part[0][48:70]
    assert np.allclose(g0[:48], exp0[:48]) and np.allclose(g1[48:], exp1[48:])
    # positions: rank 0 received 48..69 from rank 1 (value s + 0.5), rank 1 received 0..9 from rank 0 (value s)
    x0, x1 = got[0][2], got[1][2]
    assert np.allclose(x0[48:70, 0], np.arange(48, 70) + 0.5) and np.all(x0[70:] == -1.0) and got[0][3] == [(48, 70)]
    assert np.allclose(x1[0:10, 0], np.arange(0, 10)) and np.all(x1[10:40] == -1.0) and np.allclose(x1[40:48, 0], np.arange(40, 48))
    assert got[1][3] == [(0, 10), (40, 48)]
    # rebuild: everybody has everything
    for rank in range(2):
        xa = got[rank][4]
        assert np.allclose(xa[:48, 0], np.arange(48)) and np.allclose(xa[48:, 0], np.arange(48, 100) + 0.5)
