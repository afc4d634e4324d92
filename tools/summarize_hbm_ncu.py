"""Per-kernel DRAM traffic and duration of the HBM-bound kernels from an ncu metrics pass
(`HBM_REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--clock-control none --csv --log-file X.csv python tools/bench_hbm.py`).  Run here, no GPU.
usage: summarize_hbm_ncu.py <X.csv> <out.md> [hbm_peak_GBs]
Under ncu every launch is serialised and starts with cold caches: the GB/s column is
(DRAM bytes the kernel really moved) / (its duration under the profiler), i.e. achieved DRAM
throughput, not a bench value; `tools/bench_hbm.py`'s own CUDA-event timings are in r2_hbm_kernels.json."""
import csv
import sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6548.5
rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(r[0], {"name": r[4]})
    val = float(r[14].replace(",", ""))
    unit = r[13]
    scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1.0, "ms": 1e-3,
             "us": 1e-6, "ns": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
    d[r[12]] = val * scale
agg = OrderedDict()
for d in launch.values():
    if "pmb::" not in d["name"]:
        continue
    name = d["name"].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [])
    a.append((d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)))
lines = ["# HBM-bound kernels under ncu (dram__bytes_read.sum + dram__bytes_write.sum, gpu__time_duration.sum)", "",
         "source: %s; shapes of tools/bench_hbm.py (o = 27, v = 314: 0.58 GB per T2-sized tensor); the LAST launch of each"
         " kernel is listed (the first ones are warm-ups); peak = %.1f GB/s (MEASURED_PEAKS.json)" % (src, peak), "",
         "| kernel | launches | DRAM read GB | DRAM written GB | duration ms | DRAM GB/s | frac of peak |", "|---|---|---|---|---|---|---|"]
for name, ls in agg.items():
    rd, wr, dur = ls[-1]
    if dur <= 0:
        continue
    gbs = (rd + wr) / dur / 1e9
    lines.append("| `%s` | %d | %.3f | %.3f | %.3f | %.0f | %.2f |" % (name[:70], len(ls), rd / 1e9, wr / 1e9, dur * 1e3, gbs, gbs / peak))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
