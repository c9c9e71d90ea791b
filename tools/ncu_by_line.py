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
ins, cur, fn, fns = [], ("?", 0), "(kernel body)", []
for l in dis[start + 1:]:
    if l.startswith("//---"):
        break
    mf = re.match(r"^\$\S+\$_Z\d+(\w+?)(?:ILi|P|R|f|i)\S*:", l)
    if mf:
        fn = mf.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
        fns.append(fn)
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
byfn = collections.defaultdict(lambda: [0, 0, 0])
for k in range(min(len(fns), len(data))):
    g = byfn[fns[k]]
    g[0] += int(data[k][ie]); g[1] += int(data[k][ns]); g[2] += 1
print("-- by device function (noinline callees; everything else is the kernel body)")
for f, (e, sm, c) in sorted(byfn.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% inst %6.2f%% stall-smp %5d sass  %s" % (100 * e / ti, 100 * sm / max(ts, 1), c, f))
print("-- by source line")
for (f, l), (e, s, c, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    txt = src[f][l - 1].strip()[:84] if f in src and 0 < l <= len(src[f]) else ""
    print("%5.2f%% inst %5.2f%% stall-smp %4d sass %4.1f thr  %s:%d  %s" % (100 * e / ti, 100 * s / max(ts, 1), c, t / max(e, 1), f, l, txt))
