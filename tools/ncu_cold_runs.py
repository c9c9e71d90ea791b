"""The step loop of a kernel in address order, from an ncu report captured with `--set full --import-source on`: which SASS
instructions inside the loop's address span were executed (hot) and which never or rarely were (cold runs that still occupy
instruction-cache lines between hot code).  usage: python tools/ncu_cold_runs.py report.ncu-rep <kernel name regex> path/to/libcemk.so"""
import csv, os, re, subprocess, sys, tempfile
rep, kname, lib = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and kname in l][0]
ins, cur, fn = [], ("?", 0), "(body)"
for l in dis[start + 1:]:
    if l.startswith("//---"): break
    mf = re.match(r"^\$\S+\$_Z\d+(\w+?)(?:ILi|P|R|f|i)\S*:", l)
    if mf: fn = mf.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): ins.append((fn, cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[0]; b = starts[1] if len(starts) > 1 else len(rows)
print(rows[a][1][:80])
hdr, data = rows[a + 1], [r for r in rows[a + 2:b] if len(r) > 10]
ie = hdr.index("Instructions Executed")
ex = [int(r[ie]) for r in data]
n = min(len(ex), len(ins))
print("sass", len(ex), "disasm", len(ins))
HOT = 100000
# find first and last hot instruction in body -> loop span
hot_idx = [k for k in range(n) if ex[k] >= HOT and ins[k][0] == "(body)"]
lo, hi = hot_idx[0], hot_idx[-1]
print("hot span", lo, hi, "=", (hi - lo + 1) * 16 / 1024, "KB;  hot instr in span", len(hot_idx), "=", len(hot_idx) * 16 / 1024, "KB")
# cold runs inside the span
runs = []; k = lo
while k <= hi:
    if ex[k] < HOT:
        j = k
        while j <= hi and ex[j] < HOT: j += 1
        runs.append((k, j - k)); k = j
    else: k += 1
tot = sum(l for _, l in runs)
print("cold instructions inside the span:", tot, "=", tot * 16 / 1024, "KB in", len(runs), "runs")
for s0, l in sorted(runs, key=lambda t: -t[1])[:25]:
    lines = [ins[q][1][1] for q in range(s0, s0 + l) if ins[q][1][0] == "rollout_core.h"]
    med = sorted(lines)[len(lines) // 2] if lines else 0
    mx = max(ex[s0:s0 + l])
    print(f"  at {s0:6d} len {l:5d}  max exec {mx:8d}  lines {min(lines) if lines else 0}-{max(lines) if lines else 0} median {med}")
