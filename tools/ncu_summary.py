"""Summarise gpurun_out/launches.csv and gpurun_out/prof.ncu-rep into profiles/ (run in the dev container)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
g = os.path.join(ROOT, "gpurun_out")

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(g, "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list ({tag}): python bench.py --steps 2 --warmup 1 --no-cpu-baseline",
         "# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000  (cold-cache, serialised: compare SHARES)",
         f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches", "ms,launches,share_pct,kernel"]
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{t / 1e6:.4f},{n},{100 * t / tot:.2f},\"{k[:120]}\"")
open(os.path.join(out_dir, f"{tag}_launches.csv"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))

# ---- full capture
raw = subprocess.run(["ncu", "-i", os.path.join(g, "prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max"]
summ = []
for r in rows[2:]:
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = r[i] + (" " + units[i] if units[i] else "")
    summ.append(d)
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for d, r in zip(summ, rows[2:]):
    st = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:8]
    d["warps_stalled_per_issue_active"] = {n: round(v, 2) for v, n in st}
json.dump(summ, open(os.path.join(out_dir, f"{tag}_k_wave_ncu_full.json"), "w"), indent=1)
for d in summ[:2]:
    print(json.dumps(d, indent=1))
# traffic per launch of the primary-wave kernel for bench.py's roofline.traffic
def gb(s):
    v, u = s.split()
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
# the dominant kernel's counters in the form bench.py's roofline falls back to when its own ncu child cannot run
dom = [d for d in summ if "k_primary_follow" in d["Kernel Name"]] or [d for d in summ if "k_wave<0, 0, 0, 0" in d["Kernel Name"]]
out = {"source": f"profiles/{tag}_k_wave_ncu_full.json (ncu --set full, one launch each)"}
if dom:
    d0 = dom[0]
    out.update({"kernel": d0["Kernel Name"], "dram__bytes_read.sum": gb(d0["dram__bytes_read.sum"]), "dram__bytes_write.sum": gb(d0["dram__bytes_write.sum"]),
                "sm__inst_executed.sum": float(d0["smsp__inst_executed.sum"].split()[0]),
                "smsp__issue_active.avg.pct_of_peak_sustained_active": float(d0["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0]),
                "sm__warps_active.avg.pct_of_peak_sustained_active": float(d0["sm__warps_active.avg.pct_of_peak_sustained_active"].split()[0]),
                "smsp__thread_inst_executed_per_inst_executed.ratio": float(d0["smsp__thread_inst_executed_per_inst_executed.ratio"].split()[0])})
json.dump(out, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"))
print(out)
# ---- one from-scratch step, kernel by kernel (the nine launches of the full capture are the traced part of one step):
# the share of the step the roofline kernel takes, to set beside bench.py's roofline.kernel_share_of_step
step = [(d["Kernel Name"], float(d["gpu__time_duration.sum"].split()[0]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}[d["gpu__time_duration.sum"].split()[1]]) for d in summ]
tot_step = sum(t for _, t in step)
json.dump({"source": "the consecutive launches of the ncu --set full capture = every wave / projection kernel of one from-scratch step (pose update, refit and bin emission, ~0.15 ms, not captured)",
           "total_ms": round(tot_step, 4), "kernels": [{"kernel": k, "ms": round(t, 4), "share_pct": round(100 * t / tot_step, 2)} for k, t in step]},
          open(os.path.join(out_dir, f"{tag}_step_breakdown.json"), "w"), indent=1)
