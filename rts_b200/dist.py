"""Multi-GPU plumbing: one process per GPU (torchrun), scene and BVH replicated, primary rays
sharded in contiguous slabs, receiver bins reduced with one all-reduce pair per pulse.

The reference has no multi-GPU path; primary rays are independent
(/root/reference/ray_tracer.cu:151, 227-253) and the only cross-ray step is the commutative
aggregation (/root/reference/aggregation.cu:56-69), so the exchange is:
    all_reduce(SUM, float64) over bins[n][5] = {npath, sum sqrt(P), sum delay, sum phase, sum Doppler}
    all_reduce(MIN, int64)   over the smallest result-slot index per bin (orders like d_pathMatch)
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport; the bins live in the
library's own device memory and are wrapped zero-copy.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [begin, begin+count) of the primary-ray index space for `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_rays), world)
    begin = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return begin, count


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def bins_as_tensors(engine, device):
    """Zero-copy torch views of the engine's raw bin accumulators: (sums float64[n*5], mins int64[n])."""
    import torch

    sp, ns, mp, nm = engine.bins_device()
    sums = torch.as_tensor(_DevArray(sp, ns, "<f8"), device=device)
    mins = torch.as_tensor(_DevArray(mp, nm, "<i8"), device=device)
    return sums, mins


EMPTY_MIN = 0x7F7F7F7F7F7F7F7F   # what librts_b200 stores in the slot-index array of a bin nobody hit


def allreduce_bin_tensors(sums, mins, group=None):
    """The exchange step: SUM over the five fp64 accumulators, MIN over the representative slot index.
    Slot indices are < 2^63 and empty bins hold EMPTY_MIN, so MIN on the int64 view is order-preserving."""
    import torch.distributed as dist

    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    return sums, mins


_views = {}


def allreduce_bins(engine, device, group=None):
    """All-reduce the engine's bins in place over NCCL, then mark them final (myKernel2 runs at collection).
    The zero-copy views are cached per (engine, device pointers).

    Ordering: NCCL enqueues on torch's current stream of `device`, the engine on its own stream.  The helper makes the
    two the same stream (engine.set_stream, once — it waits for the engine's earlier work), so that the all-reduce runs
    behind the pulse that filled the bins (also an RTS_ASYNC one) and rts_finalise_bins behind the all-reduce."""
    import torch

    cur = torch.cuda.current_stream(device).cuda_stream or 1     # 0 is the legacy default stream: its explicit handle is cudaStreamLegacy (0x1)
    if getattr(engine, "_stream", None) != cur:
        engine.set_stream(cur)
    key = (id(engine),) + tuple(engine.bins_device())
    if key not in _views:
        _views.clear()
        _views[key] = bins_as_tensors(engine, device)
    sums, mins = _views[key]
    allreduce_bin_tensors(sums, mins, group)
    engine.finalise_bins()


def merge_bins_numpy(parts):
    """CPU statement of the same reduction for tests: list of (sums[n,5], mins[n] uint64) -> merged."""
    sums = np.sum([p[0] for p in parts], axis=0)
    mins = np.min([p[1] for p in parts], axis=0)
    return sums, mins
