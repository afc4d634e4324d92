"""Where one CCSD+DIIS sweep of the bench workload (TC-UEG 54e) spends its time, per rank.

    python tools/profile_sweep.py [cutoff] [dense] [out_prefix]                       (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29540 tools/profile_sweep.py [cutoff] [dense|virtual] [out_prefix]

Two passes over one sweep each, after two warm-up sweeps:
  1. every pmb_contract / pmb_gemv launch bracketed by CUDA events and grouped by index pattern
     (backend.enable_trace) -- which contraction costs what, and at which rate;
  2. the same sweep under torch.profiler (CUPTI): device time of EVERY kernel by name, i.e. also the
     elementwise kernels, the copies and the NCCL kernels, plus the total span of the sweep, so that
     "time in no kernel at all" (launch gaps, waits for a collective) is visible.
Diagnostic, not a bench value.  Each rank writes <out_prefix>_rank<r>.txt (default gpurun_out/sweep_profile)."""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from pymes_b200 import backend as bk, log as plog
from pymes_b200.model import ueg
from pymes_b200.solver import ccsd
from pymes_b200.integral.partition import KEYS

cutoff = float(sys.argv[1]) if len(sys.argv) > 1 else bench.CUTOFF_FOR_GPUS[1]
dense = len(sys.argv) > 2 and sys.argv[2] == "dense"
prefix = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/sweep_profile"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
plog.set_quiet(True)
no = bench.N_ELE // 2
m = ueg.UEG(bench.N_ELE, no, no, bench.RS)
m.init_single_basis(cutoff)
m.k_cutoff, m.gamma = bench.K_CUTOFF, None
fock = bench.build_fock(m, no)
virtual = () if dense else ("abcd",)
if world > 1:
    from pymes_b200 import parallel
    comm = parallel.Comm(dist.group.WORLD)
    cc = parallel.ShardedCCSD(no, comm)
    dV = parallel.build_sharded_hamiltonian(m, no, comm, bench.tc_parts(m), virtual=virtual)
else:
    cc = ccsd.CCSD(no)
    dV = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=virtual)
cc.setup(fock, dV)
for _ in range(2):
    cc.sweep()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


lines = []
# ---- pass 1: contraction trace -------------------------------------------------------------
barrier()
bk.enable_trace(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
cc.sweep()
ev1.record()
rows = bk.trace_report()
bk.enable_trace(False)
total = ev0.elapsed_time(ev1)
agg = collections.OrderedDict()
for lab, fl, ms in rows:
    a = agg.setdefault(lab, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += fl
    a[2] += ms
lines.append("rank %d of %d, nP=%d o=%d v=%d: sweep %.1f ms, %d contraction calls, %.1f ms inside them "
             "(calls on the side stream overlap: the sum can exceed the sweep)"
             % (rank, world, m.n_orb, no, m.n_orb - no, total, len(rows), sum(r[2] for r in rows)))
for lab, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    lines.append("%9.3f ms  x%-2d %7.2f TFLOP/s  %s" % (ms, n, fl / ms / 1e9 if ms > 0 else 0.0, lab))

# ---- pass 2: every kernel by name ----------------------------------------------------------------
try:
    from torch.profiler import profile, ProfilerActivity
    barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ev0.record()
        cc.sweep()
        ev1.record()
        torch.cuda.synchronize()
    span = ev0.elapsed_time(ev1)
    by_name = collections.OrderedDict()
    first, last = None, None
    for evt in prof.events():
        if evt.device_type is not None and "cuda" in str(evt.device_type).lower() and evt.device_time_total > 0:
            a = by_name.setdefault(evt.name, [0, 0.0])
            a[0] += 1
            a[1] += evt.device_time_total / 1e3
    busy = sum(v[1] for v in by_name.values())
    lines.append("")
    lines.append("kernels by name (torch.profiler / CUPTI), sweep span %.1f ms, sum of kernel times %.1f ms "
                 "(kernels on two streams overlap)" % (span, busy))
    for name, (n, ms) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append("%9.3f ms  x%-4d %s" % (ms, n, name[:150]))
except Exception as exc:          # noqa: BLE001  (CUPTI not available on the box: keep pass 1)
    lines.append("torch.profiler pass failed: %r" % (exc,))

os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
with open("%s_rank%d.txt" % (prefix, rank), "w") as fh:
    fh.write("\n".join(lines) + "\n")
if rank == 0:
    print("\n".join(lines))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
