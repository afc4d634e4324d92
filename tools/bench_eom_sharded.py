"""Batched EOM-CCSD sigma with the (ab) row blocks spread over the ranks of one box (row (e) of
SURVEY 8 for configs C4/C5): TC-UEG 54e, V_abcd never materialised (dressed V_abcd as the operator
ccsd.DressedLadder), integral blocks replicated, sigma rows sharded and all-gathered.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29520 tools/bench_eom_sharded.py [cutoff=20] [ccsd_sweeps=2] [max_batch=8]

Written at the end of round 1 after the GPU budget was spent: NOT yet run on GPUs (the same
code path passes the world-2 gloo tests in tests/test_parallel_cpu.py).  Diagnostic, not the
bench line; rank 0 prints one JSON object."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from pymes_b200 import backend as bk, log as plog, parallel
from pymes_b200.model import ueg
from pymes_b200.solver import ccsd, eom_ccsd
from pymes_b200.integral.partition import KEYS

cutoff = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
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
nv = m.n_orb - no
fock = bk.asdev(bench.build_fock(m, no))
dV = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=("abcd",))
cc = ccsd.CCSD(no)                      # replicated ground state: every rank runs the same sweeps
cc.setup(fock, dV)
for _ in range(sweeps):
    e = cc.sweep()
T1, T2 = cc._st["T1"], cc._st["T2"]
ft = cc.get_T1_dressed_fock(fock, T1, dV)
dVd = cc.get_T1_dressed_V(T1, dV, {k: None for k in eom_ccsd.V_KEYS_USED})
del dV, cc
torch.cuda.empty_cache()
shard = parallel.Shard(parallel.Comm(dist.group.WORLD), nv) if world > 1 else None
plan = eom_ccsd.SigmaPlan(no, ft, {k: dVd[k] for k in eom_ccsd.V_KEYS_USED}, T2, shard=shard)
out = {"n_gpus": world, "n_orb": m.n_orb, "n_occ": no, "n_virt": nv, "E_ccsd_after_sweeps": sum(e[:3]),
       "flops_per_vector": plan.flops_per_vector, "ladder_flops_per_vector": 2.0 * no ** 2 * nv ** 4,
       "allocated_GB": torch.cuda.memory_allocated() / 1e9, "batches": []}
torch.manual_seed(0)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
r = 1
while r <= max_batch:
    U1 = torch.randn(r, nv, no, dtype=torch.float64, device="cuda")
    U2 = torch.randn(r, nv, nv, no, no, dtype=torch.float64, device="cuda")
    if world > 1:                       # the trial vectors are replicated: same numbers on every rank
        dist.broadcast(U1, 0)
        dist.broadcast(U2, 0)
    plan.apply(U1, U2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 2
    t0.record()
    for _ in range(reps):
        S1, S2 = plan.apply(U1, U2)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    out["batches"].append({"r": r, "ms": ms, "ms_per_rhs": ms / r, "checksum": float(S2.abs().sum().item())})
    if rank == 0:
        print("r=%3d  %9.2f ms  %8.2f ms/rhs" % (r, ms, ms / r), file=sys.stderr, flush=True)
    del U1, U2, S1, S2
    r *= 2
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
