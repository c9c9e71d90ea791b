"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.
usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel mangled-name substring> [launch index] [top N]
Needs the in-tree libcemk.so that produced the report (for nvdisasm line info)."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "manipulator_mujoco_b200", "libcemk.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and kname in l][0]
ins, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
# split per kernel launch (a "Kernel Name" row starts each)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[launch]
b = starts[launch + 1] if launch + 1 < len(starts) else len(rows)
hdr, data = rows[a + 1], [r for r in rows[a + 2:b] if len(r) > 10]
ie, ns, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
print("kernel:", rows[a][1], "| sass instr:", len(data), "| disasm instr:", len(ins))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
ti = ts = 0
for k in range(min(len(ins), len(data))):
    e, s, t = int(data[k][ie]), int(data[k][ns]), int(data[k][te])
    g = agg[ins[k]]
    g[0] += e; g[1] += s; g[2] += 1; g[3] += t
    ti += e; ts += s
src = {}
for f in ("rollout_core.h", "warp_dsl.h", "cemk.cu"):
    src[f] = open(os.path.join(root, "manipulator_mujoco_b200", "csrc", f)).read().split("\n")
print("total warp-instructions %d, samples %d" % (ti, ts))
for (f, l), (e, s, c, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    txt = src[f][l - 1].strip()[:84] if f in src and 0 < l <= len(src[f]) else ""
    print("%5.2f%% inst %5.2f%% stall-smp %4d sass %4.1f thr  %s:%d  %s" % (100 * e / ti, 100 * s / max(ts, 1), c, t / max(e, 1), f, l, txt))
