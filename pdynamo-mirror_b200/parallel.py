"""Multi-GPU plumbing of the NB path: one process per GPU, spatial slabs, halo exchange.

Every rank sorts all atoms the same way (cell order, csrc/list_build.cu); rank r OWNS the contiguous slab of sorted positions that
holds its i-blocks [nblocks*r/R, nblocks*(r+1)/R) -- a spatial slab of the box.  Its pair lists reference, inside every other
rank's slab, one contiguous range of sorted positions: the halo.  Per NB call the ranks exchange exactly these ranges

    owners -> halo : positions of the halo atoms           (before the energy call; all slabs when the lists are rebuilt)
    halo -> owners : gradient contributions to halo atoms  (after it; summed into the owner's slab)

with point-to-point messages (NCCL send/recv on GPUs, gloo in the CPU tests), and all-reduce the 6 energies and dE/dM.
The reference has no counterpart (its only parallelism is an OpenMP team, NBModelABFSState.c:401)."""
import ctypes as C

import numpy as np


def block_range(nblocks, rank, nranks):
    return (nblocks * rank) // nranks, (nblocks * (rank + 1)) // nranks


def slab_range(nblocks, n, rank, nranks, tile=32):
    """Sorted positions [s0, s1) owned by `rank` (mirrors sort_and_tile() in csrc/list_build.cu)."""
    b0, b1 = block_range(nblocks, rank, nranks)
    return b0 * tile, min(n, b1 * tile)


def reduce_results(energies, dEdM, grad, group=None):
    """All-reduce (sum) the partial results of one call in place.  energies[6], dEdM[9 or 3x3] are small host arrays
    (packed into one message); grad is a torch tensor on the compute device (or None: gradients stay with their owners)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    device = grad.device if grad is not None else torch.device("cpu")
    small = torch.from_numpy(np.concatenate([np.asarray(energies, np.float64).ravel(), np.asarray(dEdM, np.float64).ravel()])).to(device)
    dist.all_reduce(small, group=group)
    if grad is not None:
        dist.all_reduce(grad, group=group)
    out = small.cpu().numpy()
    energies[:] = out[:6]
    np.asarray(dEdM).reshape(-1)[:] = out[6:15]


class SlabExchange:
    """The two halo exchanges in sorted space.  table[p, r] = [lo, hi): sorted positions rank p's lists reference inside rank r's
    slab (lo == hi: none).  Works on any torch device / backend with point-to-point support."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.table = np.zeros((world, world, 2), np.int64)
        self._recv = {}

    def set_ranges(self, mine, device):
        """mine[r] = [lo, hi) of this rank (nbb200_touched_ranges); gathers everybody's.  The own slab is never exchanged."""
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(np.asarray(mine, np.int64).reshape(self.world, 2)).to(device)
        t[self.rank] = 0
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        self.table = torch.stack(out).cpu().numpy()
        self._recv = {}

    def _run(self, ops):
        import torch.distributed as dist
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def halo_to_owners(self, gs):
        """gs[n, 3] sorted-order partial gradients: send what this rank accumulated for atoms of other slabs to their owners and
        add what the others accumulated for this rank's atoms."""
        import torch
        import torch.distributed as dist
        ops, adds = [], []
        for r in range(self.world):
            if r == self.rank:
                continue
            lo, hi = self.table[self.rank, r]
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, gs[lo:hi], r, self.group))
            lo, hi = self.table[r, self.rank]
            if hi > lo:
                key = ("g", r, int(hi - lo))
                buf = self._recv.get(key)
                if buf is None:
                    buf = self._recv[key] = torch.empty((int(hi - lo), gs.shape[1]), dtype=gs.dtype, device=gs.device)
                ops.append(dist.P2POp(dist.irecv, buf, r, self.group))
                adds.append((int(lo), int(hi), buf))
        self._run(ops)
        for lo, hi, buf in adds:
            gs[lo:hi] += buf

    def owners_to_halo(self, xs):
        """xs[n, 3] sorted-order positions, authoritative inside the own slab: send the sub-ranges the other ranks list, receive
        the own halo ranges in place."""
        import torch.distributed as dist
        ops = []
        for r in range(self.world):
            if r == self.rank:
                continue
            lo, hi = self.table[r, self.rank]
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, xs[lo:hi], r, self.group))
            lo, hi = self.table[self.rank, r]
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, xs[lo:hi], r, self.group))
        self._run(ops)
        return [(int(self.table[self.rank, r, 0]), int(self.table[self.rank, r, 1])) for r in range(self.world)
                if r != self.rank and self.table[self.rank, r, 1] > self.table[self.rank, r, 0]]

    def allgather_slabs(self, xs, slabs):
        """List rebuild: every rank needs every position.  slabs[r] = [s0, s1) of rank r; received in place."""
        import torch.distributed as dist
        ops = []
        s0, s1 = slabs[self.rank]
        for r in range(self.world):
            if r == self.rank:
                continue
            if s1 > s0:
                ops.append(dist.P2POp(dist.isend, xs[s0:s1], r, self.group))
            r0, r1 = slabs[r]
            if r1 > r0:
                ops.append(dist.P2POp(dist.irecv, xs[r0:r1], r, self.group))
        self._run(ops)


class DistributedNB:
    """Drives one NBModelABFS state per rank through the slab / halo scheme above.  x[n, 3] and g[n, 3] are device tensors in ATOM
    order on every rank; a rank keeps the positions of its own atoms current (after the first call, which takes a replicated x)
    and receives the gradient of its own atoms."""

    def __init__(self, state, n, buffer_distance, rank, world, device, group=None):
        import torch
        from . import _lib
        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.h, self.n, self.rank, self.world, self.group = state.cObject, n, rank, world, group
        self.buffac2 = float(buffer_distance) ** 2            # (listCutoff - outerCutoff) / 2, squared: CheckForUpdate's criterion
        self.L.nbb200_set_partition(self.h, rank, world)
        self.gs = torch.zeros((n, 3), dtype=torch.float64, device=device)
        self.xs = torch.zeros((n, 3), dtype=torch.float64, device=device)
        self.L.nbb200_set_sorted_gradient_buffer(self.h, C.c_void_p(self.gs.data_ptr()))
        self.exchange = SlabExchange(rank, world, group)
        self.flag = torch.zeros(1, dtype=torch.float64, device=device)
        self.small = torch.zeros(15, dtype=torch.float64, device=device)
        self.first, self.box, self.slabs = True, None, None
        self.energies, self.dEdM = np.zeros(6), np.zeros(9)
        self.updates = 0

    def _slabs(self):
        out = (C.c_long * 4)()
        self.L.nbb200_get_slab(self.h, out)
        nblocks = int(out[3])
        return [slab_range(nblocks, self.n, r, self.world) for r in range(self.world)]

    def call(self, x, box, g=None, force_rebuild=False):
        import torch.distributed as dist
        L, st = self.L, C.c_int(16)
        xp = C.c_void_p(x.data_ptr())
        box = np.ascontiguousarray(box, np.float64)
        rebuild = True
        if not self.first:
            s0, s1 = self.slabs[self.rank]
            moved = L.nbb200_max_displacement(self.h, xp, C.byref(st)) > self.buffac2
            changed = self.box is None or not np.array_equal(self.box, box)
            self.flag[0] = 1.0 if (moved or changed or force_rebuild) else 0.0
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)       # all ranks rebuild together
            rebuild = bool(self.flag.item() > 0.0)
            L.nbb200_gather_sorted(self.h, xp, s0, s1 - s0, C.c_void_p(self.xs[s0:].data_ptr()))
            if rebuild:                                      # every rank needs every position for the sort
                self.exchange.allgather_slabs(self.xs, self.slabs)
                L.nbb200_scatter_sorted(self.h, C.c_void_p(self.xs.data_ptr()), 0, self.n, xp)
            else:                                            # positions of the halo atoms only
                for lo, hi in self.exchange.owners_to_halo(self.xs):
                    L.nbb200_scatter_sorted(self.h, C.c_void_p(self.xs[lo:].data_ptr()), lo, hi - lo, xp)
        updated = L.NBModelABFS_B200_UpdateDeviceDecided(self.h, xp, self._lib.d_(box), 1 if rebuild else 0, C.byref(st))
        if st.value != 16:
            raise RuntimeError("distributed update failed: " + self._lib.last_error())
        if updated:
            self.updates += 1
            mine = (C.c_long * (2 * self.world))()
            if not L.nbb200_touched_ranges(self.h, mine):
                raise RuntimeError("touched ranges failed: " + self._lib.last_error())
            self.exchange.set_ranges(np.array(mine[:], np.int64), x.device)
            self.slabs = self._slabs()
        self.first, self.box = False, box.copy()
        L.NBModelABFS_B200_MMMMEnergySorted(self.h, self._lib.d_(self.energies), self._lib.d_(self.dEdM), C.byref(st))
        if st.value != 16:
            raise RuntimeError("distributed energy failed: " + self._lib.last_error())
        self.exchange.halo_to_owners(self.gs)
        if g is not None:
            s0, s1 = self.slabs[self.rank]
            L.nbb200_unsort_add(self.h, s0, s1 - s0, C.c_void_p(g.data_ptr()))
        self.small[:6] = self.torch.from_numpy(self.energies).to(self.small.device)
        self.small[6:] = self.torch.from_numpy(self.dEdM).to(self.small.device)
        dist.all_reduce(self.small, group=self.group)
        return updated

    def results(self):
        out = self.small.cpu().numpy()
        return out[:6].copy(), out[6:].reshape(3, 3).copy()
