"""Batched EOM-CCSD sigma (H-bar.R) throughput on the TC-UEG 54e Hamiltonian, one GPU
(BASELINE.json configs[3]/[4] at a basis whose dressed V_abcd fits one B200).
Prints ms per batch, ms per right-hand side and FP64 TFLOP/s from the plan's own per-vector
flop count, for batch sizes r = 1, 4, 16, 32(, 64).  Diagnostic, not the bench line.
usage: bench_eom.py [cutoff=13] [ccsd_sweeps=3] [max_batch=32] [virtual]
virtual: V_abcd is never materialised; its dressed form is the operator ccsd.DressedLadder (not
yet run on a GPU in round 1) and only the dressed blocks sigma reads are built."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pymes_b200 import backend as bk, log as plog
from pymes_b200.model import ueg
from pymes_b200.solver import ccsd, eom_ccsd
from pymes_b200.integral.partition import KEYS

cutoff = float(sys.argv[1]) if len(sys.argv) > 1 else 13.0
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
virtual = len(sys.argv) > 4 and sys.argv[4] == "virtual"
torch.cuda.set_device(0)
plog.set_quiet(True)
no = bench.N_ELE // 2
m = ueg.UEG(bench.N_ELE, no, no, bench.RS)
m.init_single_basis(cutoff)
m.k_cutoff, m.gamma = bench.K_CUTOFF, None
nv = m.n_orb - no
fock = bk.asdev(bench.build_fock(m, no))
dV = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=("abcd",) if virtual else ())
cc = ccsd.CCSD(no)
cc.setup(fock, dV)
for _ in range(sweeps):
    e = cc.sweep()
T1, T2 = cc._st["T1"], cc._st["T2"]
ft = cc.get_T1_dressed_fock(fock, T1, dV)
dVd = cc.get_T1_dressed_V(T1, dV, {k: None for k in eom_ccsd.V_KEYS_USED} if virtual else None)
del dV, cc
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
t0.record()
plan = eom_ccsd.SigmaPlan(no, ft, {k: dVd[k] for k in eom_ccsd.V_KEYS_USED}, T2)
t1.record()
torch.cuda.synchronize()
n_direct = sum(len(p["direct"]) for p in plan.programs.values())
n_two = sum(len(p["twostep"]) for p in plan.programs.values())
out = {"n_orb": m.n_orb, "n_occ": no, "n_virt": nv, "E_ccsd_after_sweeps": sum(e[:3]), "virtual_abcd": virtual,
       "plan_ms": t0.elapsed_time(t1), "hoist_flops": plan.hoist_flops,
       "flops_per_vector": plan.flops_per_vector, "direct_groups": n_direct, "twostep_groups": n_two,
       "ladder_flops_per_vector": 2.0 * no**2 * nv**4, "batches": []}
# with momentum-blocked products (backend.blocked_enabled) fewer flops are executed than the plan's dense
# count: one traced application (an event pair per launch, backend.enable_trace) sums what actually ran
flops_exec = plan.flops_per_vector
if bk.blocked_enabled():
    U1 = torch.randn(1, nv, no, dtype=torch.float64, device="cuda")
    U2 = torch.randn(1, nv, nv, no, no, dtype=torch.float64, device="cuda")
    plan.apply(U1, U2)
    bk.enable_trace(True)
    plan.apply(U1, U2)
    tr = bk.trace_report()
    bk.enable_trace(False)
    flops_exec = float(sum(fl for _lab, fl, _ms in tr))
    out.update(blocked_launches=sum(1 for lab, _f, _m in tr if "[momentum-blocked]" in lab),
               flops_per_vector_executed=flops_exec)
    del U1, U2
torch.manual_seed(0)
r = 1
while r <= max_batch:
    U1 = torch.randn(r, nv, no, dtype=torch.float64, device="cuda")
    U2 = torch.randn(r, nv, nv, no, no, dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.apply(U1, U2)
    before = bk.launch_count()
    reps = 3
    t0.record()
    for _ in range(reps):
        plan.apply(U1, U2)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    out["batches"].append({"r": r, "ms": ms, "ms_per_rhs": ms / r,
                           "tflops": flops_exec * r / ms / 1e9,
                           "tflops_dense_equivalent": plan.flops_per_vector * r / ms / 1e9,
                           "launches": (bk.launch_count() - before) // reps})
    print("r=%3d  %9.2f ms  %8.2f ms/rhs  %6.2f TFLOP/s" % (r, ms, ms / r, out["batches"][-1]["tflops"]),
          file=sys.stderr, flush=True)
    del U1, U2
    r *= 4 if r < 16 else 2
print(json.dumps(out))
