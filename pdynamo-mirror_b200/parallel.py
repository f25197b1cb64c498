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
    """The two halo exchanges in sorted space.  table[p, r, h] = [lo, hi): sorted positions rank p's lists reference inside the
    lower (h = 0) / upper (h = 1) half of rank r's slab (lo == hi: none).  Works on any torch device / backend with point-to-point
    support."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.table = np.zeros((world, world, 2, 2), np.int64)
        self._recv = {}

    def set_ranges(self, mine, device):
        """mine[r, h] = [lo, hi) of this rank (nbb200_touched_ranges); gathers everybody's.  The own slab is never exchanged."""
        import torch
        import torch.distributed as dist
        mine = np.asarray(mine, np.int64).reshape(self.world, 2, 2).copy()
        mine[self.rank] = 0
        t = torch.from_numpy(mine).to(device)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=device)
        if device.type == "cuda":
            dist.all_gather_into_tensor(out, t, group=self.group)
        else:                                                # gloo
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t, group=self.group)
            out = torch.stack(parts)
        self.table = out.cpu().numpy()
        self._recv = {}

    def ranges(self, p, r):
        """Non-empty ranges rank p references inside rank r's slab (adjacent halves merge into one message)."""
        (a0, b0), (a1, b1) = self.table[p, r]
        out = []
        if b0 > a0:
            out.append([int(a0), int(b0)])
        if b1 > a1:
            if out and out[-1][1] == a1:
                out[-1][1] = int(b1)
            else:
                out.append([int(a1), int(b1)])
        return out

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
            for lo, hi in self.ranges(self.rank, r):
                ops.append(dist.P2POp(dist.isend, gs[lo:hi], r, self.group))
            for lo, hi in self.ranges(r, self.rank):
                key = ("g", r, lo, hi)
                buf = self._recv.get(key)
                if buf is None:
                    buf = self._recv[key] = torch.empty((hi - lo, gs.shape[1]), dtype=gs.dtype, device=gs.device)
                ops.append(dist.P2POp(dist.irecv, buf, r, self.group))
                adds.append((lo, hi, buf))
        self._run(ops)
        for lo, hi, buf in adds:
            gs[lo:hi] += buf

    def owners_to_halo(self, xs):
        """xs[n, 3] sorted-order positions, authoritative inside the own slab: send the sub-ranges the other ranks list, receive
        the own halo ranges in place.  Returns the received ranges."""
        import torch.distributed as dist
        ops, got = [], []
        for r in range(self.world):
            if r == self.rank:
                continue
            for lo, hi in self.ranges(r, self.rank):
                ops.append(dist.P2POp(dist.isend, xs[lo:hi], r, self.group))
            for lo, hi in self.ranges(self.rank, r):
                ops.append(dist.P2POp(dist.irecv, xs[lo:hi], r, self.group))
                got.append((lo, hi))
        self._run(ops)
        return got

    def halo_atoms(self):
        """Atoms whose positions this rank receives (= whose gradient contributions it sends) per call without a rebuild."""
        return sum(hi - lo for r in range(self.world) if r != self.rank for lo, hi in self.ranges(self.rank, r))

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
    and receives the gradient of its own atoms.

    transport "peer" (default on GPUs): the ranks map each other's buffers (CUDA IPC over NVLink); the two halo exchanges AND the
    synchronisation are plain kernels of the library -- pull positions from their owners, push gradient contributions into the
    owners' accumulators with atomics, write (value, step flag) pairs into the peers' signal areas and spin (bounded) on the own one
    for the update decision and the sum of the 15 scalars.  No library collective inside a call (NCCL only hands the IPC handles
    round at set-up).  transport "p2p": the same exchanges as NCCL / gloo send/recv messages (SlabExchange) plus two all-reduces."""

    def __init__(self, state, n, buffer_distance, rank, world, device, group=None, transport=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.h, self.n, self.rank, self.world, self.group = state.cObject, n, rank, world, group
        self.buffac2 = float(buffer_distance) ** 2            # (listCutoff - outerCutoff) / 2, squared: CheckForUpdate's criterion
        device = torch.device(device)
        self.L.nbb200_set_partition(self.h, rank, world)
        if device.type == "cuda":                            # library kernels and the exchanges are ordered by one stream
            self.L.nbb200_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.transport = transport or ("peer" if device.type == "cuda" else "p2p")
        self.exchange = SlabExchange(rank, world, group)
        self.flag = torch.zeros(1, dtype=torch.float64, device=device)
        self.small = torch.zeros(15, dtype=torch.float64, device=device)
        self.small_host = torch.zeros(15, dtype=torch.float64)
        if device.type == "cuda":
            self.small_host = self.small_host.pin_memory()
        self.tab_mine = torch.zeros(4 * world, dtype=torch.int64, device=device)
        self.tab_all = torch.zeros(4 * world * world, dtype=torch.int64, device=device)
        self.tab_host = torch.zeros(4 * world * world, dtype=torch.int64).pin_memory()
        if self.transport == "peer":
            # map everybody's buffers (CUDA IPC).  The outcome is agreed collectively: if any rank cannot export / import (e.g. processes
            # in different PID namespaces), all ranks fall back to the message transport.
            buf = C.create_string_buffer(192)
            ok = bool(self.L.nbb200_peer_export(self.h, buf))
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(device)
            allh = torch.empty(192 * world, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().numpy().tobytes()
            for r in range(world):
                ok = ok and bool(self.L.nbb200_peer_import(self.h, r, allh[192 * r:192 * (r + 1)]))
            buf2 = C.create_string_buffer(128)                  # atom-order chunk buffers of call_host
            ok = ok and bool(self.L.nbb200_peer_export_chunks(self.h, buf2))
            mine2 = torch.frombuffer(bytearray(buf2.raw), dtype=torch.uint8).to(device)
            allh2 = torch.empty(128 * world, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh2, mine2, group=group)
            allh2 = allh2.cpu().numpy().tobytes()
            for r in range(world):
                ok = ok and bool(self.L.nbb200_peer_import_chunks(self.h, r, allh2[128 * r:128 * (r + 1)]))
            agreed = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)   # also: every signal area is zeroed and mapped before the first call
            if float(agreed.item()) < 1.0:
                import warnings
                warnings.warn("peer-memory transport unavailable (%s); using NCCL send/recv" % _lib.last_error())
                self.transport = "p2p"
                self.L.nbb200_set_sorted_gradient_buffer(self.h, None)
        if self.transport == "peer":
            self.gs = self.xs = None
            self.step, self.sums = 0, np.zeros(15)
            # spatial decomposition of the sort: this rank scatters / sorts / packs only the cells its slab can see (the owners publish
            # the atom indices of their slabs together with the positions, so nobody needs the whole sorted-position -> atom map)
            self.L.nbb200_set_restricted_sort(self.h, 1)
        else:
            self.gs = torch.zeros((n, 3), dtype=torch.float64, device=device)
            self.xs = torch.zeros((n, 3), dtype=torch.float64, device=device)
            self.L.nbb200_set_sorted_gradient_buffer(self.h, C.c_void_p(self.gs.data_ptr()))
        import os
        self.host_scalars = os.environ.get("NBB200_HOST_SCALARS") is not None      # the round-1 hand-over of the scalars through the host
        self.profile = None                                  # set to a dict to collect wall-clock seconds per phase (synchronising: debugging only)
        self.host_profile = None                             # the same for call_host
        self.first, self.box, self.slabs = True, None, None
        self.energies, self.dEdM = np.zeros(6), np.zeros(9)
        self.updates = 0

    def _slabs(self):
        out = (C.c_long * 4)()
        self.L.nbb200_get_slab(self.h, out)
        nblocks = int(out[3])
        return [slab_range(nblocks, self.n, r, self.world) for r in range(self.world)]

    def _tick(self, name):
        if self.profile is not None:
            import time
            self.torch.cuda.synchronize()
            now = time.perf_counter()
            self.profile[name] = self.profile.get(name, 0.0) + now - self._t
            self._t = now

    def halo_atoms(self):
        if self.transport == "peer":
            self.exchange.table[self.rank] = self.tab_mine.cpu().numpy().reshape(self.world, 2, 2)
        else:
            self.exchange.table = self.tab_all.cpu().numpy().reshape(self.world, self.world, 2, 2)
        return self.exchange.halo_atoms()

    def _decide(self, xp, box, force_rebuild, st):
        """Message transport: collective update decision (all ranks rebuild together) with an all-reduce.  force_rebuild must be the
        same on all ranks (it skips the collective)."""
        import torch.distributed as dist
        if force_rebuild:
            return True
        moved = self.L.nbb200_max_displacement(self.h, xp, C.byref(st)) > self.buffac2
        changed = self.box is None or not np.array_equal(self.box, box)
        self.flag[0] = 1.0 if (moved or changed) else 0.0
        dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(self.flag.item() > 0.0)

    def call(self, x, box, g=None, force_rebuild=False):
        import torch.distributed as dist
        L, st = self.L, C.c_int(16)
        peer = self.transport == "peer"
        if self.profile is not None:
            import time
            self.torch.cuda.synchronize()
            self._t = time.perf_counter()
        xp = C.c_void_p(x.data_ptr())
        box = np.ascontiguousarray(box, np.float64)
        rebuild = True
        changed = self.box is None or not np.array_equal(self.box, box)
        if peer:
            # begin: zero the accumulator, publish the own positions, signal; wait for everybody (also the global update decision)
            self.step += 1
            if self.first:
                L.nbb200_peer_begin(self.h, None, 0, 0)
            else:
                s0, s1 = self.slabs[self.rank]
                L.nbb200_peer_begin(self.h, xp, s0, s1 - s0)
            forced = bool(self.first or force_rebuild or changed)      # known on every rank: the decision needs no host wait
            L.nbb200_peer_signal_begin(self.h, self.step, xp, 1 if forced else 0)
            rebuild = L.nbb200_peer_wait_begin(self.h, self.step, 0 if forced else 1, C.byref(st)) > self.buffac2
            if st.value != 16:
                raise RuntimeError("distributed begin failed: " + self._lib.last_error())
            self._tick("decide")
            if not self.first:
                edges = (C.c_long * (self.world + 1))(*([sl[0] for sl in self.slabs] + [self.n]))
                L.nbb200_peer_pull_positions(self.h, C.c_void_p(self.tab_mine.data_ptr()), edges, 1 if rebuild else 0, xp)
            self._tick("positions")
        elif not self.first:
            s0, s1 = self.slabs[self.rank]
            rebuild = self._decide(xp, box, force_rebuild, st)
            self._tick("decide")
            L.nbb200_gather_sorted(self.h, xp, s0, s1 - s0, C.c_void_p(self.xs[s0:].data_ptr()))
            if rebuild:                                      # every rank needs every position for the sort
                self.exchange.allgather_slabs(self.xs, self.slabs)
                L.nbb200_scatter_sorted(self.h, C.c_void_p(self.xs.data_ptr()), 0, self.n, xp)
            else:                                            # positions of the halo atoms only
                for lo, hi in self.exchange.owners_to_halo(self.xs):
                    L.nbb200_scatter_sorted(self.h, C.c_void_p(self.xs[lo:].data_ptr()), lo, hi - lo, xp)
            self._tick("positions")
        updated = L.NBModelABFS_B200_UpdateDeviceDecided(self.h, xp, self._lib.d_(box), 1 if rebuild else 0, C.byref(st))
        if st.value != 16:
            raise RuntimeError("distributed update failed: " + self._lib.last_error())
        self._tick("update")
        if updated:
            # the halo ranges of the new lists: computed and all-gathered on the device (stream ordered, no host wait)
            self.updates += 1
            if not L.nbb200_touched_ranges_device(self.h, C.c_void_p(self.tab_mine.data_ptr())):
                raise RuntimeError("touched ranges failed: " + self._lib.last_error())
            if not peer:                                     # send/recv messages need the other ranks' tables (sizes on both sides)
                dist.all_gather_into_tensor(self.tab_all, self.tab_mine, group=self.group)
                self.tab_host.copy_(self.tab_all, non_blocking=True)
            self.slabs = self._slabs()
        self.first, self.box = False, box.copy()
        if peer:
            # the push to the owners only depends on the kernels: it is enqueued before the host waits for the energies
            L.NBModelABFS_B200_MMMMEnergySortedEnqueue(self.h, C.byref(st))
            L.nbb200_peer_push_gradients(self.h, C.c_void_p(self.tab_mine.data_ptr()))
            if not self.host_scalars:
                # the 15 scalars go from the accumulators to the peers' signal areas by kernels: no host wait inside the call at all
                # (self.energies / self.dEdM, this rank's own terms, are not updated: results() has the sums over the ranks)
                L.nbb200_peer_signal_end_device(self.h, self.step, C.byref(st))
            else:
                L.NBModelABFS_B200_MMMMEnergySortedFinish(self.h, self._lib.d_(self.energies), self._lib.d_(self.dEdM), C.byref(st))
        else:
            L.NBModelABFS_B200_MMMMEnergySorted(self.h, self._lib.d_(self.energies), self._lib.d_(self.dEdM), C.byref(st))
        if st.value != 16:
            raise RuntimeError("distributed energy failed: " + self._lib.last_error())
        self._tick("energy")
        s0, s1 = self.slabs[self.rank]
        if peer:
            if self.host_scalars:
                scal = np.concatenate([self.energies, self.dEdM])
                L.nbb200_peer_signal_end(self.h, self.step, self._lib.d_(scal))
            L.nbb200_peer_wait_end(self.h, self.step)          # on the stream: all pushes into the own slab are complete, scalars summed
            if g is not None:
                L.nbb200_unsort_add(self.h, s0, s1 - s0, C.c_void_p(g.data_ptr()))
            self._tick("scalars")
            return updated
        if updated:                                          # read after the energy call has synchronised the stream
            self.exchange.table = self.tab_host.numpy().reshape(self.world, self.world, 2, 2).copy()
            self.exchange._recv = {}
        self.exchange.halo_to_owners(self.gs)
        self._tick("gradients")
        self.small_host[:6] = self.torch.from_numpy(self.energies)
        self.small_host[6:] = self.torch.from_numpy(self.dEdM)
        self.small.copy_(self.small_host, non_blocking=True)
        dist.all_reduce(self.small, group=self.group)
        if g is not None:
            L.nbb200_unsort_add(self.h, s0, s1 - s0, C.c_void_p(g.data_ptr()))
        self._tick("scalars")
        return updated

    def call_host(self, x_host, box, g_host=None, force_rebuild=False, overwrite=False):
        """The same call for a caller that keeps coordinates and gradients in HOST arrays (x_host[n, 3], g_host[n, 3], numpy).  No row gathers on
        the host: rank r moves the CONTIGUOUS rows [n r / R, n (r + 1) / R) of the two arrays (24 n / R bytes each way, one DMA each), and the
        device redistributes over peer memory -- a rank gathers the positions of the atoms it owns from the chunk buffers of the ranks that
        uploaded them and writes the gradients of its atoms into the chunk buffers of the ranks that download them (nbb200_chunk_*).  The first
        call uploads everything once.  The rows of the rank's chunk of g_host are ACCUMULATED into (the reference's semantics) -- or, with
        overwrite, SET (System.Energy's own freshly zeroed gradient array with the NB term first: the zero fill is folded into the call, as the
        one-GPU plugin does); page-locked arrays (pdynamo_mirror_b200._lib.pinned_array, what the mirror's System allocates) travel by one DMA each
        way without a staging copy.  Returns (updated, energies[6], dEdM[3, 3])."""
        import time
        torch, L = self.torch, self.L
        if self.transport != "peer":
            raise RuntimeError("call_host needs the peer-memory transport")
        if x_host.dtype != np.float64 or not x_host.flags.c_contiguous or (g_host is not None and (g_host.dtype != np.float64 or not g_host.flags.c_contiguous)):
            raise TypeError("call_host takes C-contiguous float64 arrays")
        n = self.n
        c0, c1 = (n * self.rank) // self.world, (n * (self.rank + 1)) // self.world
        prof = self.host_profile is not None
        t0 = time.perf_counter() if prof else 0.0
        if not hasattr(self, "_xdev"):
            import os
            # threads of the host copies of this rank (read once by the library): the ranks of a box share its cores
            os.environ.setdefault("NBB200_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // max(1, self.world)))))
            self._xdev = torch.from_numpy(np.ascontiguousarray(x_host)).to(self.flag.device)
            self._hcall = 0
        else:
            self._hcall += 1
            L.nbb200_chunk_upload(self.h, C.c_void_p(x_host.ctypes.data), c0, c1 - c0)
            if prof:
                ta = time.perf_counter(); torch.cuda.synchronize(); tb = time.perf_counter()
                self.host_profile["upload: host copy"] = self.host_profile.get("upload: host copy", 0.0) + ta - t0
                self.host_profile["upload: DMA"] = self.host_profile.get("upload: DMA", 0.0) + tb - ta
            L.nbb200_chunk_signal(self.h, self._hcall, 0)
            L.nbb200_chunk_wait(self.h, self._hcall, 0)          # on the stream: every rank's rows are on its device
            L.nbb200_chunk_gather_owned(self.h, C.c_void_p(self._xdev.data_ptr()))
        if prof:
            torch.cuda.synchronize(); t1 = time.perf_counter()
        updated = self.call(self._xdev, box, None, force_rebuild)
        if prof:
            torch.cuda.synchronize(); t2 = time.perf_counter()
        self._own_count = c1 - c0
        if g_host is not None:
            self._gcall = getattr(self, "_gcall", 0) + 1
            L.nbb200_chunk_scatter_gradients(self.h)              # behind nbb200_peer_wait_end: the slab's gradients are complete
            L.nbb200_chunk_signal(self.h, self._gcall, 1)
            L.nbb200_chunk_wait(self.h, self._gcall, 1)
            if prof:
                torch.cuda.synchronize(); tc = time.perf_counter()
                self.host_profile["scatter + flags"] = self.host_profile.get("scatter + flags", 0.0) + tc - t2
            if not L.nbb200_chunk_download(self.h, C.c_void_p(g_host.ctypes.data), c0, c1 - c0, 1 if overwrite else 0):
                raise RuntimeError("distributed gradient download failed: " + self._lib.last_error())
        e, dEdM = self.results()                             # synchronises the stream
        if prof:
            t3 = time.perf_counter()
            for k, v in (("upload+gather", t1 - t0), ("call", t2 - t1), ("scatter+download+add", t3 - t2)):
                self.host_profile[k] = self.host_profile.get(k, 0.0) + v
        return updated, e, dEdM

    def results(self):
        if self.transport == "peer":
            st = C.c_int(16)
            self.L.nbb200_peer_read_sums(self.h, self._lib.d_(self.sums), C.byref(st))
            if st.value != 16:
                raise RuntimeError("distributed end failed: " + self._lib.last_error())
            return self.sums[:6].copy(), self.sums[6:].reshape(3, 3).copy()
        out = self.small.cpu().numpy()
        return out[:6].copy(), out[6:].reshape(3, 3).copy()
