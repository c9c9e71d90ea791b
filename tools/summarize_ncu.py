"""Summarise one kernel of an `ncu --set full` report into the files kept under profiles/:
    python tools/summarize_ncu.py <report.ncu-rep> <kernel name substring> <out prefix>
writes <prefix>_details.csv (the details page) and <prefix>_metrics.json (selected raw metrics + dram_traffic_bytes_per_launch)."""
import csv
import json
import subprocess
import sys

rep, kname, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(det.split("\n")) if r]
hdr = rows[0]
ki = hdr.index("Kernel Name")
with open(prefix + "_details.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(hdr)
    for r in rows[1:]:
        if len(r) > ki and kname in r[ki]:
            w.writerow(r)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.split("\n")))
h, units = rr[0], rr[1]
row = [r for r in rr[2:] if len(r) > 5 and kname in r[h.index("Kernel Name")]][0]
d = dict(zip(h, row))
u = dict(zip(h, units))
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
keep += [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
out = {k: {"unit": u.get(k, ""), "value": d[k]} for k in keep if k in d}


def num(k):
    v = float(d[k].replace(",", ""))
    unit = u.get(k, "").lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)


out["dram_traffic_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
with open(prefix + "_metrics.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: v["value"] if isinstance(v, dict) else v for k, v in out.items()}, indent=1)[:3000])
