"""Multi-GPU host logic on CPU: ray sharding and the receiver-bin exchange (SUM over the five fp64
accumulators, MIN over the representative slot index) with torch.distributed/gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rts_b200 import dist as rdist


def test_shard_ranges_partition_the_index_space():
    for n in (0, 1, 7, 100, 16777216, 10 ** 8):
        for world in (1, 2, 3, 4, 8):
            spans = [rdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        rdist.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bins, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    sums = rng.uniform(0, 1, size=(n_bins, 5))
    mins = rng.integers(0, 1 << 40, size=n_bins).astype(np.uint64)
    empty = rng.uniform(size=n_bins) < 0.5          # bins this rank never touched
    sums[empty] = 0
    mins[empty] = np.uint64(rdist.EMPTY_MIN)
    np.save(os.path.join(out_dir, f"sums{rank}.npy"), sums)
    np.save(os.path.join(out_dir, f"mins{rank}.npy"), mins)
    ts = torch.from_numpy(sums.reshape(-1).copy())
    tm = torch.from_numpy(mins.view(np.int64).copy())
    rdist.allreduce_bin_tensors(ts, tm)
    np.save(os.path.join(out_dir, f"rsums{rank}.npy"), ts.numpy().reshape(n_bins, 5))
    np.save(os.path.join(out_dir, f"rmins{rank}.npy"), tm.numpy().view(np.uint64))
    dist.destroy_process_group()


def test_bin_allreduce_gloo_world2(tmp_path):
    world, n_bins = 2, 257
    mp.spawn(_worker, args=(world, _free_port(), n_bins, str(tmp_path)), nprocs=world, join=True)
    parts = [(np.load(tmp_path / f"sums{r}.npy"), np.load(tmp_path / f"mins{r}.npy")) for r in range(world)]
    esums, emins = rdist.merge_bins_numpy(parts)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rsums{r}.npy"), esums)      # two addends: order-independent
        assert np.array_equal(np.load(tmp_path / f"rmins{r}.npy"), emins)
    assert (emins == np.uint64(rdist.EMPTY_MIN)).any()                      # bins empty on every rank stay empty
