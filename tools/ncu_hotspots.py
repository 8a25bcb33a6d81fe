"""Join the ncu SASS page of a capture (gpurun_out/prof.ncu-rep) with nvdisasm line info of the built object and
aggregate executed instructions / stall samples per source line and per function region of trace.cu.
    python tools/ncu_hotspots.py [kernel-substring] [kernel-instance-index]"""
import csv, os, re, subprocess, sys, collections, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else "k_wave<(bool)1, (bool)0, (bool)0, (bool)0>"
mangled = sys.argv[2] if len(sys.argv) > 2 else "k_waveILb1ELb0ELb0ELb0"
rep = os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rts_b200", "csrc", "trace.o")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
# instructions of the function with their (file, line)
ins = []; cur = ("?", 0); on = False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        on = mangled in l
        continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
# function regions of trace.cu by line
src = open(os.path.join(ROOT, "rts_b200", "csrc", "trace.cu")).read().splitlines()
regions = []
for i, l in enumerate(src, 1):
    m = re.match(r'(?:template.*\n)?(?:__device__ __forceinline__|__global__).*?\b(\w+)\(', l)
    if m and not l.startswith(" "): regions.append((i, m.group(1)))
def region(f, ln):
    if f != "trace.cu": return f
    name = "?"
    for s, n in regions:
        if s <= ln: name = n
    return name
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sections = []; k = None
for r in rows:
    if r and r[0] == "Kernel Name": k = [r[1], None, []]; sections.append(k); continue
    if r and r[0] == "Address": k[1] = r; continue
    if k and k[1]: k[2].append(r)
for name, hdr, body in sections:
    if pat not in name: continue
    ci = {n: i for i, n in enumerate(hdr)}
    n = min(len(body), len(ins))
    print(f"== {name}: {len(body)} SASS rows, {len(ins)} disassembled")
    tot = sum(float(b[ci["Instructions Executed"]] or 0) for b in body)
    tots = sum(float(b[ci["# Samples"]] or 0) for b in body)
    by_reg = collections.defaultdict(lambda: [0.0, 0.0, 0.0]); by_line = collections.defaultdict(lambda: [0.0, 0.0])
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    by_reg_stall = collections.defaultdict(lambda: collections.Counter())
    for b, (off, text, (f, ln)) in zip(body, ins):
        ie = float(b[ci["Instructions Executed"]] or 0); sm = float(b[ci["# Samples"]] or 0)
        rg = region(f, ln)
        by_reg[rg][0] += ie; by_reg[rg][1] += sm; by_reg[rg][2] += float(b[ci["Thread Instructions Executed"]] or 0)
        by_line[(f, ln)][0] += ie; by_line[(f, ln)][1] += sm
        for c in stall_cols:
            v = float(b[ci[c]] or 0)
            if v: by_reg_stall[rg][c] += v
    print(f"total warp-instructions {tot:.3e}, samples {tots:.0f}")
    print("-- by region: inst%  samples%  top stalls")
    for rg, (ie, sm, tie) in sorted(by_reg.items(), key=lambda x: -x[1][1]):
        top = ", ".join(f"{c[6:]}={v / max(sm,1) * 100:.0f}%" for c, v in by_reg_stall[rg].most_common(4))
        print(f"  {rg:28s} {100 * ie / tot:6.2f} {100 * sm / tots:6.2f}  thr/inst {tie / max(ie, 1):5.1f}   {top}")
    print("-- top lines by samples")
    for (f, ln), (ie, sm) in sorted(by_line.items(), key=lambda x: -x[1][1])[:28]:
        text = src[ln - 1].strip()[:90] if f == "trace.cu" and ln <= len(src) else ""
        print(f"  {f}:{ln:<5d} inst {100 * ie / tot:5.2f}%  samples {100 * sm / tots:5.2f}%  {text}")
    break
