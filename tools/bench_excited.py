"""Excited-state configurations of BASELINE.json on the TC-UEG 54e Hamiltonian:

  C4  `davidson`: EOM-CCSD, lowest n roots by block Davidson (eom_ccsd.py:46-167), batched sigma;
  C5  `feast`   : FEAST-EOM-CCSD contour (feast_eom_ccsd.py:72-181) with `nodes` quadrature points x
                  `trial` trial vectors = nodes x trial shifted linear systems, 2 real right-hand
                  sides each, advanced in lock-step through the batched sigma.

    python tools/bench_excited.py davidson [cutoff=13] [roots=10] [max_iter=40] [scalar|diagonal]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29530 tools/bench_excited.py feast [cutoff=13] [nodes=16] [trial=32] [feast_iters=1] \\
        [e_c] [e_r] [max_systems=8] [krylov=20]

With N > 1 ranks the work is dealt out over the ranks (`parallel="vectors"` / `"systems"`): every
rank holds the whole operator -- V_abcd is never materialised (its T1-dressed form is the operator
ccsd.DressedLadder), so it fits -- and applies it to its own share of the trial vectors / linear
systems; NCCL carries the all-reduce of the sigma vectors (Davidson) or of the filtered vectors Q
and H-bar Q (FEAST).  Rank 0 prints one JSON object.  Diagnostic, not the bench line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from pymes_b200 import backend as bk, log as plog, parallel
from pymes_b200.integral.partition import KEYS
from pymes_b200.model import ueg
from pymes_b200.solver import ccsd, eom_ccsd, feast_eom_ccsd


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _mem_gb():
    return torch.cuda.memory_allocated() / 1e9 if torch.cuda.is_available() else 0.0


def ground_state(cutoff, n_ele=None, delta_e=1e-9, max_sweeps=30, log=None):
    """TC-UEG Hamiltonian blocks (V_abcd generated), converged CCSD amplitudes, the dressed Fock
    matrix and the dressed blocks sigma reads.  Everything else is freed."""
    n_ele = n_ele or bench.N_ELE
    no = n_ele // 2
    m = ueg.UEG(n_ele, no, no, bench.RS)
    m.init_single_basis(cutoff)
    m.k_cutoff, m.gamma = bench.K_CUTOFF, None
    nv = m.n_orb - no
    fock = bk.asdev(bench.build_fock(m, no))
    dV = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=("abcd",))
    cc = ccsd.CCSD(no)
    cc.setup(fock, dV)
    e_last, sweeps = None, 0
    for sweeps in range(1, max_sweeps + 1):
        e = sum(cc.sweep()[:3])
        if e_last is not None and abs(e - e_last) < delta_e:
            break
        e_last = e
    if log:
        log("CCSD: %d sweeps, E_corr = %.10f" % (sweeps, e))
    T1, T2 = cc._st["T1"], cc._st["T2"]
    ft = cc.get_T1_dressed_fock(fock, T1, dV)
    dVd = cc.get_T1_dressed_V(T1, dV, {k: None for k in eom_ccsd.V_KEYS_USED})
    del dV, cc
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return dict(model=m, no=no, nv=nv, fock=ft, dV={k: dVd[k] for k in eom_ccsd.V_KEYS_USED}, T2=T2,
                e_ccsd=e, ccsd_sweeps=sweeps)


def sigma_profile(plan, no, nv, r=1):
    """Executed flops (sum over the contraction launches actually made) and time of one batched
    sigma of r right-hand sides."""
    gen = torch.Generator(device="cpu").manual_seed(1)
    U1 = torch.randn(r, nv, no, dtype=torch.float64, generator=gen).to(bk.device())
    U2 = torch.randn(r, nv, nv, no, no, dtype=torch.float64, generator=gen).to(bk.device())
    plan.apply(U1, U2)
    _sync()
    t0 = time.perf_counter()
    plan.apply(U1, U2)
    _sync()
    dt = time.perf_counter() - t0
    flops = None
    if torch.cuda.is_available():
        bk.enable_trace(True)
        plan.apply(U1, U2)
        flops = sum(fl for _lab, fl, _ms in bk.trace_report())
        bk.enable_trace(False)
    return {"r": r, "ms": dt * 1e3, "ms_per_rhs": dt * 1e3 / r, "executed_flops_per_rhs": flops / r if flops else None,
            "executed_tflops": flops / dt / 1e12 if flops else None}


def run_davidson(argv, comm, log):
    cutoff = float(argv[0]) if len(argv) > 0 else 13.0
    roots = int(argv[1]) if len(argv) > 1 else 10
    max_iter = int(argv[2]) if len(argv) > 2 else 40
    precond = argv[3] if len(argv) > 3 else "scalar"       # "scalar": the reference's; "diagonal": extension
    gs = ground_state(cutoff, log=log)
    no, nv = gs["no"], gs["nv"]
    eom = eom_ccsd.EOM_CCSD(no, n_excit=roots, comm=comm, parallel="vectors")
    eom.preconditioner = precond
    eom.max_iter = max_iter
    eom.max_rhs = 4 if nv > 300 else 10
    t0 = time.perf_counter()
    plan = eom.plan(gs["fock"], gs["dV"], gs["T2"])
    _sync()
    t_plan = time.perf_counter() - t0
    prof = sigma_profile(plan, no, nv, r=1)
    log("plan %.1f s, sigma %.1f ms/rhs, %.1f GB allocated" % (t_plan, prof["ms_per_rhs"], _mem_gb()))
    l0 = bk.launch_count()
    t0 = time.perf_counter()
    e = eom.solve(gs["fock"], gs["dV"], gs["T2"])
    _sync()
    dt = time.perf_counter() - t0
    # residuals |H u - e u| of the final Ritz vectors (the n lowest span the first n_excit slots after a
    # collapse; otherwise report the spread of the last two Ritz-value sets through e_excit)
    return {"config": "C4 EOM-CCSD Davidson, %d roots, TC-UEG 54e" % roots, "n_orb": gs["model"].n_orb, "n_occ": no,
            "n_virt": nv, "E_ccsd": gs["e_ccsd"], "ccsd_sweeps": gs["ccsd_sweeps"], "plan_seconds": t_plan,
            "preconditioner": precond, "sigma": prof, "davidson_iterations": eom.iterations, "davidson_seconds": dt,
            "converged": bool(eom.iterations < max_iter), "roots_Eh": [float(x) for x in np.sort(e)],
            "launches": bk.launch_count() - l0, "allocated_GB_peak": torch.cuda.max_memory_allocated() / 1e9
            if torch.cuda.is_available() else None}


def run_feast(argv, comm, log):
    cutoff = float(argv[0]) if len(argv) > 0 else 13.0
    nodes = int(argv[1]) if len(argv) > 1 else 16
    trial = int(argv[2]) if len(argv) > 2 else 32
    iters = int(argv[3]) if len(argv) > 3 else 1
    e_c = float(argv[4]) if len(argv) > 4 else None
    e_r = float(argv[5]) if len(argv) > 5 else None
    max_systems = int(argv[6]) if len(argv) > 6 else 8
    krylov = int(argv[7]) if len(argv) > 7 else 20
    gs = ground_state(cutoff, log=log)
    no, nv = gs["no"], gs["nv"]
    if e_c is None:
        # a window over the lowest bare excitation energies of the dressed Fock matrix
        fd = bk.tonumpy(gs["fock"]).diagonal()
        gaps = np.sort((fd[no:, None] - fd[None, :no]).ravel())
        e_c, e_r = float(0.5 * (gaps[0] + gaps[min(20, len(gaps) - 1)])), \
            float(0.5 * (gaps[min(20, len(gaps) - 1)] - gaps[0]) + 0.05)
    fe = feast_eom_ccsd.FEAST_EOM_CCSD(no, e_c=e_c, e_r=e_r, n_trial=trial, max_iter=iters, comm=comm,
                                       parallel="systems")
    fe.n_nodes = nodes
    fe.n_excit = trial                 # start with the full trial space (the reference grows it 2 at a time)
    fe.max_systems = max_systems
    fe.max_rhs = 2 * max_systems
    fe.ls_restart = krylov
    fe.ls_recycle = 0                  # no recycle space: the single cycle below is `krylov` steps long
    fe.ls_max_iter = 1                 # one GMRES cycle per system: bounded cost, residuals are reported
    np.random.seed(7)
    t0 = time.perf_counter()
    plan = fe.plan(gs["fock"], gs["dV"], gs["T2"])
    _sync()
    t_plan = time.perf_counter() - t0
    prof = sigma_profile(plan, no, nv, r=min(2 * max_systems, 16))
    log("plan %.1f s, sigma %.1f ms/rhs at r=%d, %.1f GB allocated" % (t_plan, prof["ms_per_rhs"], prof["r"], _mem_gb()))
    l0 = bk.launch_count()
    t0 = time.perf_counter()
    ev = fe.solve(gs["fock"], gs["dV"], gs["T2"])
    _sync()
    dt = time.perf_counter() - t0
    world = comm.size if comm is not None else 1
    mv = torch.tensor([float(fe.ls_matvecs), dt], dtype=torch.float64, device=bk.device())
    mv_max = mv.clone()
    if comm is not None and comm.size > 1:
        dist.all_reduce(mv[:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(mv_max, op=dist.ReduceOp.MAX)
    total_rhs = 2.0 * float(mv[0].item())
    secs = float(mv_max[1].item())
    ev = np.asarray(ev)
    inside = ev[np.abs(ev - e_c) < e_r]
    return {"config": "C5 FEAST-EOM-CCSD, %d nodes x %d trial vectors = %d shifted systems, TC-UEG 54e, "
                      "systems dealt out over %d GPUs" % (nodes, trial, nodes * trial, world),
            "n_gpus": world, "n_orb": gs["model"].n_orb, "n_occ": no, "n_virt": nv, "E_ccsd": gs["e_ccsd"],
            "e_c": e_c, "e_r": e_r, "krylov_dim": krylov, "gmres_cycles": fe.ls_max_iter, "max_systems_in_lockstep": max_systems,
            "plan_seconds": t_plan, "sigma": prof, "feast_iterations": fe.iterations, "feast_seconds_max_over_ranks": secs,
            "real_rhs_total": total_rhs, "ms_per_rhs_whole_job": secs * 1e3 / total_rhs if total_rhs else None,
            "executed_tflops_whole_job": (prof["executed_flops_per_rhs"] * total_rhs / secs / 1e12)
            if prof["executed_flops_per_rhs"] else None,
            "per_iteration_rank0": fe.timings, "eigenvalues_inside_contour": [complex(x).real for x in np.sort_complex(inside)],
            "n_eigenvalues": int(len(ev)), "launches_rank0": bk.launch_count() - l0,
            "allocated_GB_peak": torch.cuda.max_memory_allocated() / 1e9 if torch.cuda.is_available() else None}


def main(argv):
    mode = argv[0] if argv else "davidson"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1:
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = parallel.Comm(dist.group.WORLD)
    plog.set_quiet(True)

    def log(*a):
        if rank == 0:
            print(*a, file=sys.stderr, flush=True)

    bench.claim_stdout()          # libraries print to stdout too (NCCL's version banner): keep it for the JSON
    out = (run_feast if mode == "feast" else run_davidson)(argv[1:], comm, log)
    out["n_gpus"] = world
    if rank == 0:
        bench.emit(out)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main(sys.argv[1:])
