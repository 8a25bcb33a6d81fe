#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-tracing radar path on the 1M-triangle scene (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA library (librts_b200.so)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU oracle on the host cores

Workload (config C4 of BASELINE.json / SURVEY.md Appendix C): 1,000,000-triangle synthetic terrain
+ 16 moving targets (8 boxes, 8 icospheres, 4 rotating), maxRefl = 3, 1 receiver.  One step = one
pulse: per-pulse target poses (host -> device), device rigid transform + BVH refit, all bounce waves
of a (1, 4096, 4096) ray grid per GPU with fused receiver-bin aggregation, and for N > 1 the NCCL
all-reduce of the bins (rays of a (1, 4096*N, 4096) launch are dealt round-robin to the ranks, scene and
BVH replicated: weak scaling; --scaling strong keeps the (1, 4096, 4096) launch and splits it N ways,
--scaling pulse gives every rank whole pulses, p mod N).  Pulses advance every step, so the movers really move.

value  : K steps timed without reading the bins back (inputs resident, CUDA events on the engine's stream); every pulse
         traced from scratch (RTS_NO_REUSE); the figure with the library's between-pulse reuse on is in `temporal_reuse`
e2e    : the same steps through the C-ABI with host buffers, plus the device->host read of the bins
sustained: the `value` loop again for at least --sustain seconds (the K-step legs last only tens of milliseconds)
roofline: the longest single kernel of a step (k_primary_follow: the projected primary wave's shading pass with every first
         reflection traced in place).  `bound` names what binds it; `achieved`/`frac`/`traffic` are its measured
         DRAM bytes (ncu counters of one launch, taken by a child run of this very script at the end, or failing that
         the committed capture under profiles/) against the measured HBM peak; `algorithmic` keeps SURVEY.md §8(d)'s byte
         model; `issue` is the kernel's instruction issue rate against the SMs' peak.
cpu_baseline: the oracle (BVH mode, OpenMP) on a strided sample of the same pulse
"""
from __future__ import annotations

import argparse
import csv
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
WORKLOAD = ("C4: 1,000,000-triangle terrain + 16 moving targets, per-pulse pose update + BVH refit, "
            f"(1,{N_GRID},{N_GRID}) rays per GPU per pulse, maxRefl=3, 1 Rx, fused bins")
NCU_METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,"
               "l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,"
               "sm__inst_executed.sum,sm__inst_executed.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,"
               "sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,"
               "smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,launch__registers_per_thread,"
               "launch__occupancy_limit_registers,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed")


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


def host_threads() -> int:
    """Host cores this process may use (torch.distributed.run exports OMP_NUM_THREADS=1: not what a CPU baseline wants)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- the ncu child: three pulses of the workload, nothing else; the parent reads the counters of one launch ----------
def run_ncu_child(args):
    from rts_b200 import lib as L
    ms = build_scene(1)
    eng = L.Engine(0)
    eng.set_targets(ms.base)
    for p in range(3):
        eng.set_poses(*ms.poses(p))
        eng.trace(ms.spec_for(p), L.RTS_OUT_BINS | L.RTS_NO_REUSE)
    print("child done", eng.stats()["segments"], flush=True)
    eng.close()


def ncu_counters(kernel_regex: str, skip: int):
    """One launch of the dominant kernel under ncu (counters only; no number printed under ncu is a bench value)."""
    from shutil import which
    if not which("ncu"):
        return None, "ncu not on PATH"
    out_csv = os.path.join(ROOT, "gpurun_out", "bench_ncu_child.csv")
    os.makedirs(os.path.dirname(out_csv), exist_ok=True)
    cmd = ["ncu", "--metrics", NCU_METRICS, "--clock-control", "none", "-k", f"regex:{kernel_regex}", "-s", str(skip), "-c", "1",
           "--csv", "--log-file", out_csv, sys.executable, os.path.abspath(__file__), "--ncu-child"]
    try:
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    except Exception as ex:  # noqa: BLE001
        return None, f"ncu child failed to run: {ex}"
    if r.returncode != 0 or not os.path.exists(out_csv):
        return None, f"ncu child rc={r.returncode}: {(r.stderr or r.stdout)[-200:]}"
    vals, kernel = {}, None
    with open(out_csv, newline="") as f:
        rows = [row for row in csv.reader(f) if len(row) > 10]
    if not rows:
        return None, "ncu child wrote no rows"
    hdr = rows[0]
    try:
        i_name, i_val, i_k = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    except ValueError:
        return None, "unexpected ncu csv header"
    for row in rows[1:]:
        try:
            vals[row[i_name]] = float(row[i_val].replace(",", ""))
            kernel = row[i_k]
        except ValueError:
            pass
    if "dram__bytes_read.sum" not in vals:
        return None, "ncu child: counters missing"
    # ncu reports byte metrics in the unit it likes; ask for the unit column too
    i_unit = hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1.0, "inst": 1.0}
    for row in rows[1:]:
        if row[i_name] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            vals[row[i_name]] = float(row[i_val].replace(",", "")) * scale.get(row[i_unit], 1.0)
        if row[i_name] == "gpu__time_duration.sum":
            vals[row[i_name]] = float(row[i_val].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row[i_unit], 1e-6)
    vals["kernel"] = kernel
    return vals, "ncu child run of this script (one launch, counters only)"


def run_ours(args):
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

    mode = args.scaling if world > 1 else "weak"
    ms = build_scene(world if mode == "weak" else 1)
    eng = L.Engine(local)
    eng.set_targets(ms.base)
    info = eng.bvh_info()
    log(f"[bench] rank {rank}: BVH built in {info.ms_build:.2f} ms ({info.n_nodes} nodes)")
    # all of the engine's work and torch's events on one stream
    # above the library's low-priority side stream (direction pass and static footprints of the next pulse)
    stream = torch.cuda.Stream(dev, priority=int(os.environ.get("RTS_BENCH_STREAM_PRIORITY", "-1")))
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    # weak:   the launch grid is (1, 4096*N, 4096) — N x denser in azimuth — and rank r traces rays r, r+N, r+2N, ... of it,
    #         so every rank sees the N = 1 ray density, coherence and workload; bins all-reduced every pulse
    # strong: the (1, 4096, 4096) launch itself dealt round-robin to the ranks; bins all-reduced every pulse
    # pulse:  every rank traces whole pulses (p mod N == rank) of the (1, 4096, 4096) launch; no exchange inside the
    #         timed region (pulses are independent, ray_tracer.cpp:843); a step = N pulses, one per rank
    if mode == "pulse":
        begin, count, stride = 0, 0, 1
    else:
        begin, count, stride = rank, 0, world
    n_mine = (ms.spec.rays - begin + stride - 1) // stride
    reduce_bins = (world > 1 and mode != "pulse") or args.force_exchange
    # the bins' exchange: two kernels over peer memory (comm.cu; handles swapped through torch.distributed once), or the
    # pair of NCCL all-reduces (--exchange nccl, and the fallback where CUDA IPC is not available)
    px = None
    exchange = "none"
    if reduce_bins and world == 1:      # measurement aid: the exchange path on one GPU (its own block only)
        class _Solo:
            def allreduce_bins(self):
                eng.comm_allreduce_bins()

            def close(self):
                eng.comm_destroy()
        eng.comm_create(0, 1, 1 << 16)
        px, exchange = _Solo(), "peer (world 1)"
    elif reduce_bins:
        exchange = "nccl"
        if args.exchange == "peer":
            try:
                px = rdist.PeerExchange(eng, dev, max_bins=1 << 16)
                exchange = "peer"
            except Exception as ex:  # noqa: BLE001
                log(f"[bench] rank {rank}: peer-memory exchange unavailable ({ex}); using NCCL all-reduces")
            ok = torch.tensor([1 if px is not None else 0], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            if int(ok) == 0 and px is not None:
                px.close(); px = None; exchange = "nccl"

    # host inputs of a pulse — the rts_pose array and the rts_pulse struct with its host arrays — are prepared ahead of the
    # timed regions (a host simulator has them before it calls in); the timed region moves them to the device
    prepared = {}

    def prepare(pulse):
        if pulse not in prepared:
            spec = ms.spec_for(pulse)
            spec.ray_begin, spec.ray_count, spec.ray_stride = begin, count, stride
            prepared[pulse] = (L.Engine.pack_poses(*ms.poses(pulse)), eng.prepare(spec))
        return prepared[pulse]

    def step(pulse, read_back, reuse=False):
        # everything below is enqueued on one stream; nothing waits on the host unless the bins are read back.
        # reuse=False (the headline): RTS_NO_REUSE, every pulse is traced from scratch — ray generation, primary
        # visibility and every bounce wave; nothing computed for an earlier pulse is used.
        if mode == "pulse":
            pulse = pulse * world + rank
        (poses, n_poses), cpulse = prepare(pulse)
        eng.set_poses_packed(poses, n_poses)                 # H2D poses + device transform + refit of the movers
        eng.trace_prepared(cpulse, L.RTS_OUT_BINS | L.RTS_ASYNC | (0 if reuse else L.RTS_NO_REUSE) | (L.RTS_NO_FINALISE if reduce_bins else 0))
        if px is not None:
            px.allreduce_bins()
        elif reduce_bins:
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
        for i in range(k):
            prepare((k0 + i) * world + rank if mode == "pulse" else k0 + i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.kernel_launches()
        t0 = time.perf_counter()
        e0.record(stream)
        waves, split, follow, segs, caps, d2h = [], [], [], 0, 0, 0
        for i in range(k):
            st, bins = step(k0 + i, read_back, reuse)
            if st is not None:
                waves.append(eng.wave_profile())
                split.append(eng.split_profile())
                follow.append(eng.follow_profile())
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
        return dict(ms=float(t[0]), own_ms=ms_dev, wall_ms=float(t[1]), waves=waves, split=split, follow=follow, segments=segs, captured=caps, d2h=d2h,
                    launches=eng.kernel_launches() - launches0)

    def timed_pipelined(k0, k):
        """The e2e loop of a host that keeps one pulse in flight: step i's inputs go to the device and pulse i is enqueued,
        then pulse i-1's bins are read (rts_get_bins_previous: waits for pulse i-1's own read-back only).  Every step's
        inputs are copied in and every step's bins are read back inside the timed region, the last one behind the loop."""
        for i in range(k):
            prepare((k0 + i) * world + rank if mode == "pulse" else k0 + i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_read = 0
        for i in range(k):
            step(k0 + i, False)
            if i > 0:
                n_read += 1 if len(eng.bins_previous()) >= 0 else 0
        n_read += 1 if len(eng.bins()) >= 0 else 0
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        assert n_read == k
        return dict(ms=float(t[0]))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # nvidia-smi needs ~0.5 s to start sampling: begin before the warm-up
    for w in range(args.warmup):
        step(w, True)
    k0 = args.warmup
    def comm_mark(tag, prev=[None]):
        # peer exchange: mean wait for the slowest rank / kernel time per exchange since the last mark (stderr, every rank)
        if px is None:
            return
        c = eng.comm_stats()
        if prev[0] is not None and c["exchanges"] > prev[0]["exchanges"]:
            n = c["exchanges"] - prev[0]["exchanges"]
            w = (c["wait_us"] * c["exchanges"] - prev[0]["wait_us"] * prev[0]["exchanges"]) / n
            k = (c["kernel_us"] * c["exchanges"] - prev[0]["kernel_us"] * prev[0]["exchanges"]) / n
            log(f"[bench] rank {rank}: {tag}: {n} exchanges, waited {w:.1f} us for the slowest rank, reduce kernel {k:.1f} us on average")
        prev[0] = c

    comm_mark("warm-up")
    r_dev = timed(k0, args.steps, read_back=False)
    comm_mark("value leg")
    r_e2e = timed(k0, args.steps, read_back=True)
    comm_mark("e2e leg")
    r_pipe = None
    try:
        r_pipe = timed_pipelined(k0, args.steps)
    except Exception as ex:  # noqa: BLE001  (e.g. more than 256 non-empty bins: the pipelined read needs the eager block)
        log(f"[bench] rank {rank}: pipelined e2e leg not run ({ex})")
    ok_pipe = torch.tensor([1 if r_pipe is not None else 0], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ok_pipe, op=torch.distributed.ReduceOp.MIN)
    if int(ok_pipe) == 0:
        r_pipe = None
    # the sustained leg: the same loop for at least --sustain seconds (same K on every rank)
    r_sus = None
    if args.sustain > 0 and not args.quick:
        k_sus = max(args.steps, int(args.sustain * 1e3 / max(1e-3, r_dev["ms"] / args.steps)) + 1)
        r_sus = timed(k0, k_sus, read_back=False)
        r_sus["steps"] = k_sus
    # the same steps with what the library keeps between pulses of one launch geometry switched on (static primary hits,
    # static first-reflection hits, ray directions: raster.cuh / coherent.cuh) — reported beside the headline, not as it
    r_dev_reuse = r_e2e_reuse = None
    if not args.quick:
        for w in range(2):
            step(k0 + w, True, reuse=True)
        r_dev_reuse = timed(k0, args.steps, read_back=False, reuse=True)
        r_e2e_reuse = timed(k0, args.steps, read_back=True, reuse=True)
    clk = clocks.stop() if rank == 0 else None

    pulses_per_step = world if mode == "pulse" else 1
    rays_per_step_total = ms.spec.rays * pulses_per_step      # all ranks together
    value = rays_per_step_total * args.steps / (r_dev["ms"] * 1e-3) / 1e6
    e2e = rays_per_step_total * args.steps / (r_e2e["ms"] * 1e-3) / 1e6

    # The dominant kernel, rank 0's launches, timed live with the CUDA events the engine records on its stream.  Wave 0 is a
    # group of kernels (projected primary wave: directions, footprints, then k_primary_follow = shading of the primary hits
    # with every first reflection traced in place, follow.cuh) — k_primary_follow is the longest single kernel of a step and
    # has its own pair of events.  When it did not run (option no_follow), the second wave holds the longest kernel:
    # k_traverse (split.cuh) or the fused k_wave.
    n_w = max((len(w) for w in r_e2e["waves"]), default=0)
    nst = max(1, len(r_e2e["waves"]))
    per_wave_ms = [sum(w[i][0] for w in r_e2e["waves"] if len(w) > i) / nst for i in range(n_w)]
    per_wave_seg = [sum(w[i][1] for w in r_e2e["waves"] if len(w) > i) / nst for i in range(n_w)]
    if world > 1:
        log(f"[bench] rank {rank}: waves ms/step {[round(x, 4) for x in per_wave_ms]}, own value-leg device time {r_dev['own_ms'] / args.steps:.4f} ms/step")
    trav_ms = sum(s[0] for s in r_e2e["split"]) / nst
    shade_ms = sum(s[1] for s in r_e2e["split"]) / nst
    follow_ms = sum(r_e2e["follow"]) / nst
    followed = follow_ms > 0
    dom = 0 if followed else (max(range(1, n_w), key=lambda i: per_wave_ms[i]) if n_w > 1 else 0)
    split_on = not followed and dom == 1 and trav_ms > 0
    if followed:
        kernel_name, ncu_regex = "k_primary_follow<RECORDS=false> (primary shading pass + first reflections in place, follow.cuh)", "k_primary_follow"
        avg_ms = follow_ms
        # segments of that launch that walk the BVH: the first reflections (wave 0 counts primaries + followed segments)
        avg_seg = per_wave_seg[0] - n_mine
    elif split_on:
        kernel_name, ncu_regex, avg_ms, avg_seg = "k_traverse<COUNT=false> (second wave, split.cuh)", "k_traverse", trav_ms, per_wave_seg[dom]
    else:
        kernel_name = f"k_wave<PRIMARY=false,RECORDS=false,COUNT=false,CHAIN=false> (wave {dom})" if dom else "primary wave"
        ncu_regex, avg_ms, avg_seg = "k_wave<\\(bool\\)0", per_wave_ms[dom], per_wave_seg[dom]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # 1632 B per segment that walks the tree; the fused kernel also streams 32 B per primary ray (hit word +
    # direction) and gathers an 80-byte triangle record per primary hit (= per followed segment)
    alg_bytes = avg_seg * B_SEG_1M + (n_mine * 32.0 + avg_seg * 80.0 if followed else 0.0)
    alg_gbs = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roof = {"bound": "issue/latency", "kernel": kernel_name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src, "ms_per_launch": round(avg_ms, 4), "segments_per_launch": int(avg_seg),
            "kernel_share_of_step": round(avg_ms / (r_e2e["ms"] / args.steps), 4),
            "algorithmic": {"bytes_per_segment": B_SEG_1M, "bytes_per_launch": int(alg_bytes), "gbs": round(alg_gbs, 1),
                            "frac_of_hbm_peak": round(alg_gbs / peak, 4),
                            "note": "SURVEY.md §8(d) byte model (every node and triangle of a ray's path counted as fetched from DRAM); "
                                    "neighbouring rays share nodes through L1/L2, so this is not a DRAM rate and may exceed 1"},
            "waves_ms_per_step": [round(x, 4) for x in per_wave_ms],
            "second_wave_kernels_ms": {"k_traverse": round(trav_ms, 4), "k_shade_wave": round(shade_ms, 4)},
            "k_primary_follow_ms": round(follow_ms, 4),
            "all_waves_ms_per_step": round(sum(sum(x[0] for x in w) for w in r_e2e["waves"]) / nst, 4)}
    if rank == 0 and world == 1:
        counters, source = (None, "skipped (--quick / --no-ncu)")
        if not (args.quick or args.no_ncu):
            # the engine must let go of the GPU's profiler-visible state? no: ncu profiles the child process only
            counters, source = ncu_counters(ncu_regex, 2 if followed else 3)
            if counters is None:
                log(f"[bench] {source}")
        if counters is None:
            prof = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(prof):
                try:
                    counters = json.load(open(prof))
                    source = f"profiles/r02_traffic.json (committed ncu capture of the same command; in-run capture unavailable: {source})"
                except Exception:
                    counters = None
        if counters:
            dram = counters.get("dram__bytes_read.sum", 0.0) + counters.get("dram__bytes_write.sum", 0.0)
            roof["traffic"] = int(dram)
            roof["achieved"] = round(dram / (avg_ms * 1e-3) / 1e9, 1)
            roof["frac"] = round(roof["achieved"] / peak, 4)
            sm_clock = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
            inst = counters.get("sm__inst_executed.sum", 0.0)
            issue_peak = 4.0 * 148 * float(sm_clock) * 1e6         # one warp instruction per SM sub-partition per cycle
            roof["issue"] = {"achieved": round(inst / (avg_ms * 1e-3) / 1e9, 1), "peak": round(issue_peak / 1e9, 1), "unit": "G warp-inst/s",
                             "frac": round(inst / (avg_ms * 1e-3) / issue_peak, 4) if avg_ms > 0 else None,
                             "peak_source": f"4 schedulers x 148 SMs x {sm_clock} MHz (median SM clock of this run)"}
            roof["measured"] = {"source": source, "kernel": counters.get("kernel"),
                                "l2_bytes": int(counters.get("lts__t_sectors.sum", 0.0) * 32),
                                "l2_hit_pct": counters.get("lts__t_sector_hit_rate.pct"), "l1_hit_pct": counters.get("l1tex__t_sector_hit_rate.pct"),
                                "l1_data_pipe_pct": counters.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                                "dram_pct": counters.get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                                "l2_pct": counters.get("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                                "issue_active_pct": counters.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                "inst_executed_pct": counters.get("sm__inst_executed.avg.pct_of_peak_sustained_elapsed"),
                                "warps_active_pct": counters.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                "lanes_per_inst": counters.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                                "local_ld_inst": counters.get("smsp__inst_executed_op_local_ld.sum"),
                                "local_st_inst": counters.get("smsp__inst_executed_op_local_st.sum"),
                                "registers": counters.get("launch__registers_per_thread"),
                                "ncu_ms_cold": counters.get("gpu__time_duration.sum")}
        else:
            roof["measured"] = {"source": source}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        cpu = cpu_baseline(ms, args.cpu_stride or 1)

    if rank == 0:
        seg_per_ray = r_e2e["segments"] / (n_mine * args.steps)
        h2d = len(ms.base) * 112 + len(ms.base) * 24 + 64   # poses + target velocities + receiver
        scaling_note = {"weak": f", launch grid (1,{N_GRID * world},{N_GRID}) ray-sharded round-robin, bins reduced across the GPUs every pulse",
                        "strong": f", the (1,{N_GRID},{N_GRID}) launch ray-sharded round-robin over {world} GPUs, bins reduced across the GPUs every pulse",
                        "pulse": f", pulse-sharded: rank r traces pulses p = r mod {world} in full, no exchange; a step = {world} pulses"}[mode] if world > 1 else ""
        out = {
            "metric": "Mrays/s (3-bounce, 1M-tri scene)", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r_dev["ms"] / args.steps, 4),
            "higher_is_better": True, "scaling": "weak" if mode in ("weak", "pulse") else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD + scaling_note,
                       "triangles": int(sum(len(t.tris) for t in ms.base)), "rays_per_step": int(rays_per_step_total),
                       "segments_per_ray": round(seg_per_ray, 4), "captured_per_step": int(r_e2e["captured"] / args.steps),
                       "l2_policy": "per-step working set (BVH 64 MB + triangle records 81 MB + 0.4 GB of ray directions and 0.13 GB of hit words written and re-read) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"{'pulse' if mode == 'pulse' else 'ray'}-shard x{world}", "sharding": mode, "bin_exchange": exchange},
            "e2e": {"value": round(e2e, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(r_e2e["d2h"] / args.steps),
                    "ms_per_step": round(r_e2e["ms"] / args.steps, 4),
                    "note": "every step waits for its own bins before the next pulse is enqueued (the reference's host loop); 'pipelined' = the same copies with one pulse in flight: pulse i is enqueued, then pulse i-1's bins are read (rts_get_bins_previous)",
                    "pipelined": None if r_pipe is None else {"value": round(rays_per_step_total * args.steps / (r_pipe["ms"] * 1e-3) / 1e6, 2), "unit": "Mrays/s",
                                                               "ms_per_step": round(r_pipe["ms"] / args.steps, 4)}},
            "gpu_launches": int(r_dev["launches"]),
            "roofline": roof,
            "clocks": clk,
            "msegments_per_s": round(r_e2e["segments"] * (world if mode != "pulse" else world) / (r_dev["ms"] * 1e-3) / 1e6, 2),
        }
        if r_sus is not None:
            out["sustained"] = {"value": round(rays_per_step_total * r_sus["steps"] / (r_sus["ms"] * 1e-3) / 1e6, 2), "unit": "Mrays/s",
                                "steps": r_sus["steps"], "seconds": round(r_sus["ms"] * 1e-3, 3), "ms_per_step": round(r_sus["ms"] / r_sus["steps"], 4)}
        if r_dev_reuse is not None:
            nr = max(1, len(r_e2e_reuse["waves"]))
            out["temporal_reuse"] = {
                "value": round(rays_per_step_total * args.steps / (r_dev_reuse["ms"] * 1e-3) / 1e6, 2),
                "e2e": round(rays_per_step_total * args.steps / (r_e2e_reuse["ms"] * 1e-3) / 1e6, 2), "unit": "Mrays/s",
                "ms_per_step": round(r_dev_reuse["ms"] / args.steps, 4),
                "waves_ms_per_step": [round(sum(w[i][0] for w in r_e2e_reuse["waves"] if len(w) > i) / nr, 4)
                                      for i in range(max((len(w) for w in r_e2e_reuse["waves"]), default=0))],
                "note": "same steps with hits of the static geometry kept between pulses of one launch geometry (bit-identical results); "
                        "value / e2e above trace every pulse from scratch (RTS_NO_REUSE)"}
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if px is not None:
        barrier()              # nobody unmaps a block a peer may still be reading
        px.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    eng.close()


def cpu_baseline(ms, stride, pulse=3):
    """The oracle (port of the reference math, BVH mode, OpenMP over all host cores) on every
    `stride`-th primary ray of one pulse of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    cores = O.set_num_threads(host_threads())
    spec = ms.spec_for(pulse)
    spec.ray_begin, spec.ray_count, spec.ray_stride = 0, N_GRID * N_GRID, stride
    world = ms.world_targets(pulse)
    t0 = time.time()
    bins, st = O.trace_bins(world, spec, use_bvh=True)
    wall = time.time() - t0
    trace_s = st["ms_trace"] * 1e-3
    return {"value": round(st["primary_rays"] / trace_s / 1e6, 4), "unit": "Mrays/s", "cores": int(cores),
            "kind": "port", "sample": f"every {stride}th primary ray of pulse {pulse} ({st['primary_rays']} rays, {st['segments']} segments); "
                                      f"trace+aggregate {trace_s:.2f}s, oracle BVH build {st['ms_update'] * 1e-3:.2f}s excluded, wall {wall:.1f}s"}


def run_reference(args):
    """--impl reference: the reference's CPU-runnable restatement (the oracle port; the reference itself needs
    OptiX + SOARS and its sources-plus-shim build is exhaustive-search, single-threaded) on ALL the host cores this
    process may use — also under torch.distributed.run, which exports OMP_NUM_THREADS=1 to its workers.  The scene
    comes from the oracle's own mesh helpers: this process never maps librts_b200.so."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores_want = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores_want)          # before liboracle (and its OpenMP runtime) is loaded
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    from rts_b200 import scenes
    scenes.set_helpers(O)
    cores = O.set_num_threads(cores_want)
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
        log(f"[bench] reference arm: {rate / 1e6:.2f} Mrays/s on the probe step with {cores} threads -> every {stride}th ray per step")
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
    sample = f"every {stride}th primary ray of a (1,{N_GRID},{N_GRID}) pulse per step ({total_rays // max(1, args.steps)} rays/step), oracle BVH build excluded"
    out = {"impl": "reference", "metric": "Mrays/s (3-bounce, 1M-tri scene)", "value": round(v, 4), "unit": "Mrays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sum(per) / len(per), 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "triangles": int(sum(len(t.tris) for t in ms.base)), "rays_per_step": N_GRID * N_GRID,
                      "reference_arm": f"CPU oracle port (OpenMP, {cores} threads, oracle BVH rebuilt for every pulse's poses), bounded sample: " + sample},
           "cpu_baseline": {"value": round(v, 4), "unit": "Mrays/s", "cores": int(cores), "kind": "port", "sample": sample},
           "e2e": {"value": round(v, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong", "pulse"], help="how N > 1 GPUs share the work (see the module docstring)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: bins reduced by the library's peer-memory kernels or by NCCL all-reduces")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained leg (0 = off)")
    ap.add_argument("--cpu-stride", type=int, default=0, help="oracle sample: every n-th primary ray (0 = 1 for cpu_baseline; for --impl reference 0 = chosen so that the run takes about 100 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="do not spawn the ncu child for the roofline counters")
    ap.add_argument("--quick", action="store_true", help="A/B runs: only the value and e2e legs")
    ap.add_argument("--force-exchange", action="store_true", help="N = 1: run the peer-memory exchange path on the one GPU (measures its fixed cost)")
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.ncu_child:
        run_ncu_child(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
