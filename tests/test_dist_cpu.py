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


def _sparse_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    pool = np.arange(1, 400, dtype=np.int64) * 1_000_003          # key space shared by the ranks; each holds a subset
    mine = np.sort(rng.choice(pool, size=int(rng.integers(0 if rank else 120, 200)), replace=False))
    sums = rng.uniform(0.5, 2.0, size=(len(mine), 5))
    mins = rng.integers(0, 1 << 40, size=len(mine)).astype(np.int64)
    np.savez(os.path.join(out_dir, f"in{rank}.npz"), keys=mine, sums=sums, mins=mins)
    u, us, um = rdist.exchange_sparse(torch.from_numpy(mine), torch.from_numpy(sums), torch.from_numpy(mins))
    np.savez(os.path.join(out_dir, f"out{rank}.npz"), keys=u.numpy(), sums=us.numpy(), mins=um.numpy())
    dist.destroy_process_group()


def test_sparse_bin_exchange_gloo_world2(tmp_path):
    """exchange_sparse: all-gather of keys -> sorted union -> SUM / MIN all-reduce of the compact arrays; every rank ends up
    with the same merged table, equal to a host-side merge of the inputs."""
    world = 2
    mp.spawn(_sparse_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ins = [np.load(tmp_path / f"in{r}.npz") for r in range(world)]
    want = {}
    for d in ins:
        for k, s, m in zip(d["keys"], d["sums"], d["mins"]):
            a = want.setdefault(int(k), [np.zeros(5), 1 << 62])
            a[0] = a[0] + s
            a[1] = min(a[1], int(m))
    keys = np.array(sorted(want), dtype=np.int64)
    for r in range(world):
        o = np.load(tmp_path / f"out{r}.npz")
        assert np.array_equal(o["keys"], keys)
        assert np.allclose(o["sums"], np.stack([want[int(k)][0] for k in keys]), rtol=0, atol=1e-15)
        assert np.array_equal(o["mins"], np.array([want[int(k)][1] for k in keys], dtype=np.int64))
    # the single-process statement used by the GPU tests agrees
    u, us, um = rdist.merge_compact([(torch.from_numpy(d["keys"]), torch.from_numpy(d["sums"]), torch.from_numpy(d["mins"])) for d in ins])
    assert np.array_equal(u.numpy(), keys) and np.array_equal(um.numpy(), np.load(tmp_path / "out0.npz")["mins"])
    assert np.allclose(us.numpy(), np.load(tmp_path / "out0.npz")["sums"], rtol=0, atol=1e-15)
