"""Worker of tests/test_gpu_peer_exchange.py, one process per GPU under torch.distributed.run: a ray-sharded pulse whose bins
are reduced (a) by the library's peer-memory exchange over CUDA IPC and (b) by the NCCL all-reduce pair; rank 0 also
traces the whole launch alone.  All three must agree: keys, counts and representative slots exactly, sums to 1e-9."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rts_b200 import dist as rdist, lib as L, scenes  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ms = scenes.terrain_scene(n=512, cells_x=96, cells_y=48, movers=6, n_rx=3)
    eng = L.Engine(local)
    eng.set_targets(ms.base)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    px = rdist.PeerExchange(eng, dev, max_bins=1 << 16)
    out = {"rank": rank, "ok": True, "pulses": 0, "bins": 0, "max_rel": 0.0}
    fields = ["sum_sqrt_power", "sum_delay", "sum_phase", "sum_doppler", "power", "delay", "phase", "doppler"]
    for pulse in range(6):
        eng.set_poses(*ms.poses(pulse))
        spec = ms.spec_for(pulse)
        spec.ray_begin, spec.ray_count, spec.ray_stride = rank, 0, world
        flags = L.RTS_OUT_BINS | L.RTS_ASYNC | L.RTS_NO_FINALISE | L.RTS_NO_REUSE
        eng.trace(spec, flags)
        px.allreduce_bins()
        a = eng.bins().copy()
        eng.trace(spec, flags)
        rdist.allreduce_bins(eng, dev)
        b = eng.bins().copy()
        whole = ms.spec_for(pulse)
        eng.trace(whole, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        c = eng.bins().copy()
        for got in (a, b):
            same = len(got) == len(c) and all(np.array_equal(got[f], c[f]) for f in ("rx", "path", "npath", "min_slot", "direct"))
            out["ok"] = out["ok"] and bool(same)
            if same:
                for f in fields:
                    d = np.abs(got[f] - c[f]) / np.maximum(np.abs(c[f]), 1e-300)
                    out["max_rel"] = max(out["max_rel"], float(d.max()) if len(d) else 0.0)
        # the peer exchange sums in rank order on every GPU: all ranks hold the same bits
        mine = torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy()).to(dev)
        ref = mine.clone()
        dist.broadcast(ref, 0)
        out["ok"] = out["ok"] and bool(torch.equal(mine, ref))
        out["pulses"] += 1
        out["bins"] = int(len(c))
    out["ok"] = out["ok"] and out["max_rel"] <= 1e-9 and out["bins"] >= 2
    # device time of the two exchanges on an otherwise idle, synchronised pair of GPUs (skew excluded): informational
    for name, fn in (("peer_ms", px.allreduce_bins), ("nccl_ms", lambda: rdist.allreduce_bins(eng, dev))):
        tot = 0.0
        for rep in range(12):
            eng.trace(spec, flags)
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if rep >= 2:
                tot += e0.elapsed_time(e1)
        out[name] = round(tot / 10, 4)
    dist.barrier()
    px.close()
    eng.close()
    print("PEER_WORKER " + json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
