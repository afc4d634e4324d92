"""Summarise an `ncu --set full` report of contract_kernel into profiles/ (run here, no GPU).
usage: summarize_ncu.py <report.ncu-rep> <tag> [algorithmic_flops] [algorithmic_bytes]"""
import csv
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
flops = float(sys.argv[3]) if len(sys.argv) > 3 else None
abytes = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = []
for vals in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            d[h] = (v + (" " + u if u else "")).strip()
    out.append(d)
os.makedirs("profiles", exist_ok=True)
lines = ["# ncu --set full summary: %s" % tag, "", "source report: %s (kept in gpurun_out/, not tracked)" % rep, ""]
for n, d in enumerate(out):
    lines.append("## launch %d" % n)
    for k in KEEP:
        if k in d:
            lines.append("- `%s`: %s" % (k, d[k]))
    def num(key):
        v = d.get(key, "0").split()
        x = float(v[0].replace(",", ""))
        unit = v[1] if len(v) > 1 else ""
        return x * {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(unit, 1)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    dur = num("gpu__time_duration.sum")
    lines.append("- DRAM traffic per launch: %.3f GB" % (traffic / 1e9))
    if flops:
        lines.append("- algorithmic flops per launch: %.4e -> %.2f TFLOP/s under the profiler "
                     "(cold, serialised; not a bench value)" % (flops, flops / dur / 1e12))
    if abytes:
        lines.append("- algorithmic bytes per launch: %.3f GB (traffic / algorithmic = %.2f)"
                     % (abytes / 1e9, traffic / abytes))
    lines.append("")
    if n == 0:
        json.dump({"tag": tag, "dram_bytes_per_launch": traffic, "duration_s_under_ncu": dur,
                   "algorithmic_flops": flops, "algorithmic_bytes": abytes},
                  open("profiles/%s_traffic.json" % tag, "w"), indent=1)
open("profiles/%s_ncu.md" % tag, "w").write("\n".join(lines))
print("\n".join(lines))
