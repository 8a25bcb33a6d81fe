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


class PeerExchange:
    """The bins' exchange over peer memory instead of NCCL (include/rts_b200.h: rts_comm_*; csrc/comm.cu): every rank's
    exchange block is mapped into every other rank's address space through CUDA IPC — the 64-byte handles travel over
    torch.distributed once — and from then on a pulse's exchange is two small kernels on the engine's stream that
    publish this rank's accumulators, wait for the peers' and reduce all of them in rank order over NVLink.

        px = PeerExchange(engine, device, max_bins)      # collective: every rank of the group
        engine.trace(spec, RTS_OUT_BINS | RTS_NO_FINALISE | RTS_ASYNC); px.allreduce_bins()   # every rank, every pulse
    """

    def __init__(self, engine, device, max_bins: int, group=None):
        import torch
        import torch.distributed as dist

        self.engine = engine
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        engine.comm_create(rank, world, max_bins)
        mine = torch.frombuffer(bytearray(engine.comm_ipc_handle()), dtype=torch.uint8).to(device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        engine.comm_connect_ipc(b"".join(bytes(h.cpu().numpy().tobytes()) for h in handles))
        dist.barrier(group=group)          # nobody publishes before everyone has mapped everyone

    def allreduce_bins(self):
        self.engine.comm_allreduce_bins()

    def close(self):
        self.engine.comm_destroy()


def exchange_sparse(keys, sums, mins, group=None):
    """The exchange step for sparse bins ("hash -> sorted key table + all-gather of keys"): every rank
    holds its occupied bins as compact arrays keys int64[n_r] (distinct), sums float64[n_r,5], mins int64[n_r].
    all_gather of the counts and of the (padded) keys -> sorted union U of all ranks' keys -> each rank scatters its
    rows to their positions in U -> all_reduce SUM over [|U|,5] and MIN over [|U|].  Returns (U, sums_U, mins_U), equal
    on all ranks.  Works on any backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = keys.device
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    n_max = int(max(int(c) for c in counts))
    if n_max == 0:
        return keys[:0], sums.reshape(-1, 5)[:0], mins[:0]
    pad = torch.full((n_max,), -1, dtype=torch.int64, device=dev)      # keys are < 2^63: -1 never occurs
    pad[: keys.numel()] = keys
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    union = torch.unique(torch.cat(gathered))                          # sorted
    union = union[union >= 0]
    at = torch.searchsorted(union, keys)
    u_sums = torch.zeros((union.numel(), 5), dtype=torch.float64, device=dev)
    u_mins = torch.full((union.numel(),), EMPTY_MIN, dtype=torch.int64, device=dev)
    u_sums[at] = sums.reshape(-1, 5)
    u_mins[at] = mins
    dist.all_reduce(u_sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(u_mins, op=dist.ReduceOp.MIN, group=group)
    return union, u_sums, u_mins


def merge_compact(parts):
    """Single-process statement of exchange_sparse for tests: list of (keys, sums[n,5], mins) tensors -> merged."""
    import torch

    keys = torch.cat([p[0] for p in parts])
    union, inv = torch.unique(keys, return_inverse=True)
    u_sums = torch.zeros((union.numel(), 5), dtype=torch.float64, device=keys.device)
    u_sums.index_add_(0, inv, torch.cat([p[1].reshape(-1, 5) for p in parts]))
    u_mins = torch.full((union.numel(),), EMPTY_MIN, dtype=torch.int64, device=keys.device)
    u_mins.scatter_reduce_(0, inv, torch.cat([p[2] for p in parts]), reduce="amin")
    return union, u_sums, u_mins


def compact_bins_as_tensors(engine, device):
    """Copies (torch-owned) of the engine's occupied sparse bins: (keys int64[n], sums float64[n,5], mins int64[n])."""
    import torch

    kp, sp, mp, n = engine.bins_compact_device()
    if n == 0:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return z, torch.zeros((0, 5), dtype=torch.float64, device=device), z.clone()
    keys = torch.as_tensor(_DevArray(kp, n, "<i8"), device=device).clone()
    sums = torch.as_tensor(_DevArray(sp, n * 5, "<f8"), device=device).clone().reshape(n, 5)
    mins = torch.as_tensor(_DevArray(mp, n, "<i8"), device=device).clone()
    return keys, sums, mins


def allreduce_bins_sparse(engine, device, group=None):
    """allreduce_bins for pulses whose bins live in the sparse table (rts_bins_compact_device / rts_bins_load_compact)."""
    import torch

    cur = torch.cuda.current_stream(device).cuda_stream or 1
    if getattr(engine, "_stream", None) != cur:
        engine.set_stream(cur)
    keys, sums, mins = compact_bins_as_tensors(engine, device)
    union, u_sums, u_mins = exchange_sparse(keys, sums, mins, group)
    u_sums = u_sums.contiguous()
    engine.bins_load_compact(union.data_ptr(), u_sums.data_ptr(), u_mins.data_ptr(), union.numel())
    engine.finalise_bins()
    return union.numel()


def merge_bins_numpy(parts):
    """CPU statement of the same reduction for tests: list of (sums[n,5], mins[n] uint64) -> merged."""
    sums = np.sum([p[0] for p in parts], axis=0)
    mins = np.min([p[1] for p in parts], axis=0)
    return sums, mins
