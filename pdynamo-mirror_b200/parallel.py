"""Multi-GPU plumbing of the NB path: one process per GPU, i-blocks split into contiguous slabs,
one exchange step per call -- the sum-reduction of gradients, energies and dE/dM (NCCL on GPUs, gloo in CPU tests).

The slab arithmetic mirrors sort_and_tile() in csrc/list_build.cu: rank r owns i-blocks [nblocks*r/R, nblocks*(r+1)/R)."""
import numpy as np


def block_range(nblocks, rank, nranks):
    return (nblocks * rank) // nranks, (nblocks * (rank + 1)) // nranks


def reduce_results(energies, dEdM, grad, group=None):
    """All-reduce (sum) the partial results of one call in place.  energies[6], dEdM[9 or 3x3] are small host arrays
    (packed into one message); grad is a torch tensor on the compute device (or None)."""
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
