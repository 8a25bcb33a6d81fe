#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-tracing radar path on the 1M-triangle scene (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA library (librts_b200.so)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU oracle on the host cores

Workload (config C4 of BASELINE.json / SURVEY.md Appendix C): 1,000,000-triangle synthetic terrain
+ 16 moving targets (8 boxes, 8 icospheres, 4 rotating), maxRefl = 3, 1 receiver.  One step = one
pulse: per-pulse target poses (host -> device), device rigid transform + BVH refit, all bounce waves
of a (1, 4096, 4096) ray grid per GPU with fused receiver-bin aggregation, and for N > 1 the NCCL
all-reduce of the bins (rays of a (1, 4096*N, 4096) launch are dealt round-robin to the ranks, scene and
BVH replicated: weak scaling).  Pulses advance every step, so the movers really move.

value  : steps timed without reading the bins back (inputs resident, CUDA events on the engine's stream); every pulse
         traced from scratch (RTS_NO_REUSE); the figure with the library's between-pulse reuse on is in `temporal_reuse`
e2e    : the same steps through the C-ABI with host buffers, plus the device->host read of the bins
roofline: the longest single kernel of a step (k_wave, second wave), algorithmic bytes per SURVEY.md §8(d)
cpu_baseline: the oracle (BVH mode, OpenMP) on a strided sample of the same pulse
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 4096
B_SEG_1M = 1632.0          # algorithmic bytes per traced segment, T = 1M (SURVEY.md §8d / BASELINE.md §3)
B_CAPTURE = 40.0           # five fp64 bin updates per captured ray
WORKLOAD = ("C4: 1,000,000-triangle terrain + 16 moving targets, per-pulse pose update + BVH refit, "
            f"(1,{N_GRID},{N_GRID}) rays per GPU per pulse, maxRefl=3, 1 Rx, fused bins")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": None, "reasons": [], "samples": len(sm)}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                try:
                    out["sm_max_mhz"] = int(float(r[1]))
                except ValueError:
                    pass
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active") and n not in out["reasons"]:
                        out["reasons"].append(n)
        return out


def build_scene(world):
    from rts_b200 import scenes
    t0 = time.time()
    ms = scenes.terrain_scene(n=N_GRID * world, n_rx=1, nz=N_GRID)
    log(f"[bench] scene generated in {time.time() - t0:.1f}s: {sum(len(t.tris) for t in ms.base)} triangles, {len(ms.base)} targets")
    return ms


def run_ours(args):
    import numpy as np
    import torch
    from rts_b200 import dist as rdist
    from rts_b200 import lib as L

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus:
        log(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}: using WORLD_SIZE")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    ms = build_scene(world)
    eng = L.Engine(local)
    eng.set_targets(ms.base)
    info = eng.bvh_info()
    log(f"[bench] rank {rank}: BVH built in {info.ms_build:.2f} ms ({info.n_nodes} nodes)")
    # all of the engine's work and torch's events on one stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    # weak scaling: the launch grid is (1, 4096*N, 4096) — N x denser in azimuth — and rank r traces rays
    # r, r+N, r+2N, ... of it, so every rank sees the N = 1 ray density, coherence and workload
    begin, count, stride = rank, 0, world
    n_mine = (ms.spec.rays - begin + stride - 1) // stride

    def step(pulse, read_back, reuse=False):
        # everything below is enqueued on one stream; nothing waits on the host unless the bins are read back.
        # reuse=False (the headline): RTS_NO_REUSE, every pulse is traced from scratch — ray generation, primary
        # visibility and every bounce wave; nothing computed for an earlier pulse is used.
        eng.set_poses(*ms.poses(pulse))                      # H2D poses + device transform + refit of the movers
        spec = ms.spec_for(pulse)
        spec.ray_begin, spec.ray_count, spec.ray_stride = begin, count, stride
        eng.trace(spec, L.RTS_OUT_BINS | L.RTS_ASYNC | (0 if reuse else L.RTS_NO_REUSE) | (L.RTS_NO_FINALISE if world > 1 else 0))
        if world > 1:
            rdist.allreduce_bins(eng, dev)
        if read_back:
            bins = eng.bins()                                # D2H (waits for the pulse)
            return eng.stats(), bins
        return None, None

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(k0, k, read_back, reuse=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.kernel_launches()
        t0 = time.perf_counter()
        e0.record(stream)
        waves, segs, caps, d2h = [], 0, 0, 0
        for i in range(k):
            st, bins = step(k0 + i, read_back, reuse)
            if st is not None:
                waves.append(eng.wave_profile())
                segs += st["segments"]
                caps += st["captured"]
                # device -> pinned host per step: the emitted-bin block (256 x sizeof(rts_bin)), its count, the
                # counters and per-wave segment counts of the read-back block, and the SAH cost of the refit
                d2h += 256 * bins.dtype.itemsize + 4 + 80 + 256 + 8
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms_dev = e0.elapsed_time(e1)
        t = torch.tensor([ms_dev, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return dict(ms=float(t[0]), wall_ms=float(t[1]), waves=waves, segments=segs, captured=caps, d2h=d2h,
                    launches=eng.kernel_launches() - launches0)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # nvidia-smi needs ~0.5 s to start sampling: begin before the warm-up
    for w in range(args.warmup):
        step(w, True)
    k0 = args.warmup
    r_dev = timed(k0, args.steps, read_back=False)
    r_e2e = timed(k0, args.steps, read_back=True)
    # the same steps with what the library keeps between pulses of one launch geometry switched on (static primary hits,
    # static first-reflection hits, ray directions: raster.cuh / coherent.cuh) — reported beside the headline, not as it
    for w in range(2):
        step(k0 + w, True, reuse=True)
    r_dev_reuse = timed(k0, args.steps, read_back=False, reuse=True)
    r_e2e_reuse = timed(k0, args.steps, read_back=True, reuse=True)
    clk = clocks.stop() if rank == 0 else None

    rays_per_step_total = ms.spec.rays                       # all ranks together
    value = rays_per_step_total * args.steps / (r_dev["ms"] * 1e-3) / 1e6
    e2e = rays_per_step_total * args.steps / (r_e2e["ms"] * 1e-3) / 1e6

    # roofline of the dominant kernel, rank 0's launches, measured live with the CUDA events the engine records around
    # every wave; read in the e2e leg, where each step's events are collected.  Wave 0 is a group of kernels (projected
    # primary wave: directions, footprints, shading); every later wave is one launch of k_wave — the second wave
    # (first reflection, ~13.5M rays) is the longest single kernel of a step.
    n_w = max((len(w) for w in r_e2e["waves"]), default=0)
    per_wave_ms = [sum(w[i][0] for w in r_e2e["waves"] if len(w) > i) / max(1, len(r_e2e["waves"])) for i in range(n_w)]
    per_wave_seg = [sum(w[i][1] for w in r_e2e["waves"] if len(w) > i) / max(1, len(r_e2e["waves"])) for i in range(n_w)]
    dom = max(range(1, n_w), key=lambda i: per_wave_ms[i]) if n_w > 1 else 0
    all_ms = [sum(x[0] for x in w) for w in r_e2e["waves"]]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    avg_ms, avg_seg = per_wave_ms[dom], per_wave_seg[dom]
    achieved = (avg_seg * B_SEG_1M) / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": f"k_wave<PRIMARY=false,RECORDS=false,COUNT=false,CHAIN=false> (wave {dom})" if dom else "primary wave",
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
            "bytes_per_segment": B_SEG_1M, "segments_per_launch": int(avg_seg), "ms_per_launch": round(avg_ms, 4),
            "kernel_share_of_step": round(avg_ms / (r_e2e["ms"] / args.steps), 4),
            "waves_ms_per_step": [round(x, 4) for x in per_wave_ms],
            "all_waves_ms_per_step": round(sum(all_ms) / max(1, len(all_ms)), 4)}
    if rank == 0:
        # the same numerator against measured L2 bandwidth, because the nodes and triangle records a
        # wave touches are served by L1/L2, not by HBM (the measured DRAM bytes are in `traffic`)
        l2 = eng.probe_read_bandwidth(48 << 20, 100)
        hbm_rd = eng.probe_read_bandwidth(2 << 30, 3)
        roof["l2"] = {"peak": round(l2, 1), "frac": round(achieved / l2, 4), "unit": "GB/s", "hbm_read_same_kernel": round(hbm_rd, 1),
                      "peak_source": "measured in this run (rts_probe_read_bandwidth): 128-bit L1-bypassing loads over an L2-resident 48 MB buffer; "
                                     "hbm_read_same_kernel = the same kernel over 2 GB"}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("k_wave_later_dram_bytes_per_launch" if dom else "k_wave_primary_dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(ms, args.cpu_stride or 1)

    if rank == 0:
        seg_per_ray = r_e2e["segments"] / (n_mine * args.steps)
        h2d = len(ms.base) * 112 + len(ms.base) * 24 + 64   # poses + target velocities + receiver
        out = {
            "metric": "Mrays/s (3-bounce, 1M-tri scene)", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r_dev["ms"] / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD
                                   + (f", launch grid (1,{N_GRID * world},{N_GRID}) ray-sharded round-robin, NCCL all-reduce of bins" if world > 1 else ""),
                       "triangles": int(sum(len(t.tris) for t in ms.base)), "rays_per_step": int(rays_per_step_total),
                       "segments_per_ray": round(seg_per_ray, 4), "captured_per_step": int(r_e2e["captured"] / args.steps),
                       "l2_policy": "per-step working set (BVH 64 MB + triangle records 81 MB + 2.4 GB of ray queues written and re-read) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"ray-shard x{world}"},
            "e2e": {"value": round(e2e, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(r_e2e["d2h"] / args.steps),
                    "ms_per_step": round(r_e2e["ms"] / args.steps, 4)},
            "gpu_launches": int(r_dev["launches"]),
            "roofline": roof,
            "clocks": clk,
            "msegments_per_s": round(r_e2e["segments"] * world / (r_dev["ms"] * 1e-3) / 1e6, 2),
            "temporal_reuse": {"value": round(rays_per_step_total * args.steps / (r_dev_reuse["ms"] * 1e-3) / 1e6, 2),
                               "e2e": round(rays_per_step_total * args.steps / (r_e2e_reuse["ms"] * 1e-3) / 1e6, 2), "unit": "Mrays/s",
                               "ms_per_step": round(r_dev_reuse["ms"] / args.steps, 4),
                               "waves_ms_per_step": [round(sum(w[i][0] for w in r_e2e_reuse["waves"] if len(w) > i) / max(1, len(r_e2e_reuse["waves"])), 4)
                                                     for i in range(max((len(w) for w in r_e2e_reuse["waves"]), default=0))],
                               "note": "same steps with hits of the static geometry kept between pulses of one launch geometry (bit-identical results); "
                                       "value / e2e above trace every pulse from scratch (RTS_NO_REUSE)"},
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    eng.close()


def cpu_baseline(ms, stride, pulse=3):
    """The oracle (port of the reference math, BVH mode, OpenMP over all host cores) on every
    `stride`-th primary ray of one pulse of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    spec = ms.spec_for(pulse)
    spec.ray_begin, spec.ray_count, spec.ray_stride = 0, N_GRID * N_GRID, stride
    world = ms.world_targets(pulse)
    t0 = time.time()
    bins, st = O.trace_bins(world, spec, use_bvh=True)
    wall = time.time() - t0
    trace_s = st["ms_trace"] * 1e-3
    return {"value": round(st["primary_rays"] / trace_s / 1e6, 4), "unit": "Mrays/s", "cores": int(O.oracle().orc_num_threads()),
            "kind": "port", "sample": f"every {stride}th primary ray of pulse {pulse} ({st['primary_rays']} rays, {st['segments']} segments); "
                                      f"trace+aggregate {trace_s:.2f}s, oracle BVH build {st['ms_update'] * 1e-3:.2f}s excluded, wall {wall:.1f}s"}


def run_reference(args):
    """--impl reference: the reference's CPU-runnable restatement (the oracle port; the reference itself needs
    OptiX + SOARS and its sources-plus-shim build is exhaustive-search, single-threaded) on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    ms = build_scene(1)
    stride = args.cpu_stride
    if not stride:
        # bounded sample: a probe step at stride 16 gives the host's rate; the stride of the run is then chosen so that
        # all warm-up + timed steps together take about 100 s of oracle time, whatever K, W and the core count are
        spec = ms.spec_for(0)
        spec.ray_begin, spec.ray_count, spec.ray_stride = 0, N_GRID * N_GRID, 16
        _, st = O.trace_bins(ms.world_targets(0), spec, use_bvh=True)
        rate = st["primary_rays"] / (st["ms_trace"] * 1e-3)
        want = N_GRID * N_GRID * (args.warmup + args.steps) / (rate * 100.0)
        stride = int(min(64, max(2, -(-want // 1))))
        log(f"[bench] reference arm: {rate / 1e6:.2f} Mrays/s on the probe step -> every {stride}th ray per step")
    total_rays, total_s, per = 0, 0.0, []
    for i in range(args.warmup + args.steps):
        pulse = i
        spec = ms.spec_for(pulse)
        spec.ray_begin, spec.ray_count, spec.ray_stride = 0, N_GRID * N_GRID, stride
        bins, st = O.trace_bins(ms.world_targets(pulse), spec, use_bvh=True)
        if i >= args.warmup:
            total_rays += st["primary_rays"]
            total_s += st["ms_trace"] * 1e-3
            per.append(st["ms_trace"])
    v = total_rays / total_s / 1e6
    cores = int(O.oracle().orc_num_threads())
    sample = f"every {stride}th primary ray of a (1,{N_GRID},{N_GRID}) pulse per step ({total_rays // max(1, args.steps)} rays/step), oracle BVH build excluded"
    out = {"impl": "reference", "metric": "Mrays/s (3-bounce, 1M-tri scene)", "value": round(v, 4), "unit": "Mrays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sum(per) / len(per), 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "triangles": int(sum(len(t.tris) for t in ms.base)), "rays_per_step": N_GRID * N_GRID,
                      "reference_arm": "CPU oracle port (OpenMP, oracle BVH rebuilt for every pulse's poses), bounded sample: " + sample},
           "cpu_baseline": {"value": round(v, 4), "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(v, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-stride", type=int, default=0, help="oracle sample: every n-th primary ray (0 = 1 for cpu_baseline; for --impl reference 0 = chosen so that the run takes about 100 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
