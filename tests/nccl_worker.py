"""Multi-rank (NCCL) parity worker: launched by tests/test_gpu_parity.py::test_two_rank_nccl_parity as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port P tests/nccl_worker.py [out.json]

Every rank checks its own results against the CPU oracle (test infrastructure) and exits non-zero on
a mismatch.  Cases:
  1. ShardedCCSD on the non-hermitian LiH FCIDUMP (dense V_pqrs sliced per rank), sweep by sweep;
  2. ShardedCCSD on TC-UEG 14e / 57 plane waves, every rank generating ITS rows of V_abcd and of the
     o.v^3 blocks on the device, sweep by sweep against the oracle-built Hamiltonian;
  3. batched EOM-CCSD sigma at the benchmark's o = 27 (TC-UEG 54e / 65 plane waves): row-sharded
     sigma == single-GPU sigma == oracle, with the dressed V_abcd as a never-materialised operator;
  4. Davidson with the new vectors dealt out over the ranks (parallel="vectors") -> golden roots, and a
     seeded FEAST contour with the systems dealt out (parallel="systems") == the single-rank run.
Tolerances: energies 1e-10 Eh, amplitudes / sigma 1e-9 relative, roots 1e-8 Eh."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from oracle import cc_oracle as oc, ueg_oracle as uo
    from pymes_b200 import backend as bk, log, parallel
    from pymes_b200.integral.partition import KEYS, part_2_body_int
    from pymes_b200.model import ueg
    from pymes_b200.solver import ccsd, eom_ccsd, feast_eom_ccsd
    log.set_quiet(True)
    comm = parallel.Comm(dist.group.WORLD)
    report = {"world": world}
    launches0 = bk.launch_count()

    # ---- 1. LiH transcorrelated FCIDUMP --------------------------------------------------
    g = np.load(os.path.join(ROOT, "tests", "golden", "mol_LiH_tc.npz"))
    no = int(g["n_elec"]) // 2
    fock, V = g["fock"], g["V"]
    for is_dcsd in (False, True):
        cc = parallel.ShardedCCSD(no, comm, is_dcsd=is_dcsd)
        cc.setup(fock, V.copy())
        dV = oc.partition(no, V)
        eps_i, eps_a = fock.diagonal()[:no].copy(), fock.diagonal()[no:].copy()
        _, T2 = oc.mp2(eps_i, eps_a, dV["ijab"], dV["abij"])
        T1 = np.zeros((fock.shape[0] - no, no))
        d1, d2 = oc.denominators(eps_i, eps_a)
        mixer = oc.DIIS(6)
        worst = [0.0, 0.0]
        for sweep in range(8):
            T1, T2, e, _ = oc.ccsd_sweep(no, fock, dV, T1, T2, d1, d2, mixer, is_dcsd=is_dcsd)
            got = cc.sweep()
            worst[0] = max(worst[0], abs(sum(got[:3]) - sum(e)))
            worst[1] = max(worst[1], rel(cc._st["T2"].cpu().numpy(), T2), rel(cc._st["T1"].cpu().numpy(), T1))
        assert worst[0] < 1e-10 and worst[1] < 1e-9, ("LiH_tc", is_dcsd, worst)
        report["LiH_tc_%s" % ("dcsd" if is_dcsd else "ccsd")] = worst

    # ---- 2. TC-UEG 14e, rows generated per rank --------------------------------------------
    no = 7
    mo = uo.UEG(14, 1.0).init_single_basis(5.0)
    mo.k_cutoff, mo.gamma = bench.K_CUTOFF, None
    fock, V = mo.tc_hamiltonian(no)
    m = ueg.UEG(14, no, no, 1.0)
    m.init_single_basis(5.0)
    m.k_cutoff, m.gamma = bench.K_CUTOFF, None
    fock_dev = bench.build_fock(m, no)
    assert rel(fock_dev, fock) < 1e-11
    dVl = parallel.build_sharded_hamiltonian(m, no, comm, bench.tc_parts(m), virtual=("abcd",))
    cc = parallel.ShardedCCSD(no, comm)
    cc.setup(fock_dev, dVl)
    dV = oc.partition(no, V)
    eps_i, eps_a = fock.diagonal()[:no].copy(), fock.diagonal()[no:].copy()
    _, T2 = oc.mp2(eps_i, eps_a, dV["ijab"], dV["abij"])
    T1 = np.zeros((m.n_orb - no, no))
    d1, d2 = oc.denominators(eps_i, eps_a)
    mixer = oc.DIIS(6)
    worst = [0.0, 0.0]
    for sweep in range(5):
        T1, T2, e, _ = oc.ccsd_sweep(no, fock, dV, T1, T2, d1, d2, mixer)
        got = cc.sweep()
        worst[0] = max(worst[0], abs(sum(got[:3]) - sum(e)))
        worst[1] = max(worst[1], rel(cc._st["T2"].cpu().numpy(), T2), rel(cc._st["T1"].cpu().numpy(), T1))
    assert worst[0] < 1e-10 and worst[1] < 1e-9, ("TC-UEG 14e", worst)
    report["tc_ueg_14e_ccsd"] = worst
    del cc, dVl

    # ---- 3. EOM sigma at o = 27 --------------------------------------------------------------
    no = bench.N_ELE // 2
    prob = bench.CpuProblem(6.0)
    nv = prob.n_orb - no
    m = ueg.UEG(bench.N_ELE, no, no, bench.RS)
    m.init_single_basis(6.0)
    m.k_cutoff, m.gamma = bench.K_CUTOFF, None
    fock_dev = bk.asdev(bench.build_fock(m, no))
    dVd_in = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=("abcd",))
    cc = ccsd.CCSD(no)
    cc.setup(fock_dev, dVd_in)
    for _ in range(3):
        cc.sweep()
    T1d, T2d = cc._st["T1"], cc._st["T2"]
    ft = cc.get_T1_dressed_fock(fock_dev, T1d, dVd_in)
    dVd = cc.get_T1_dressed_V(T1d, dVd_in, {k: None for k in eom_ccsd.V_KEYS_USED})
    dVd = {k: dVd[k] for k in eom_ccsd.V_KEYS_USED}
    rng = np.random.default_rng(4)
    U1 = rng.standard_normal((3, nv, no))
    U2 = rng.standard_normal((3, nv, nv, no, no))
    U1d, U2d = bk.asdev(U1), bk.asdev(U2)
    single = eom_ccsd.SigmaPlan(no, ft, dVd, T2d)
    sharded = eom_ccsd.SigmaPlan(no, ft, dVd, T2d, shard=parallel.Shard(comm, nv))
    S1a, S2a = single.apply(U1d, U2d)
    S1b, S2b = sharded.apply(U1d, U2d)
    T1h, T2h = T1d.cpu().numpy(), T2d.cpu().numpy()
    fto = oc.dressed_fock(no, prob.fock, T1h, prob.dV)
    dVo = oc.dressed_V(T1h, prob.dV)
    worst = 0.0
    for k in range(3):
        s1 = oc.eom_sigma_singles(no, fto, dVo, U1[k], U2[k], T2h)
        s2 = oc.eom_sigma_doubles(no, fto, dVo, U1[k], U2[k], T2h)
        for got1, got2 in ((S1a, S2a), (S1b, S2b)):
            worst = max(worst, rel(got1[k].cpu().numpy(), s1), rel(got2[k].cpu().numpy(), s2))
    assert worst < 1e-9, ("EOM sigma o=27", worst)
    assert rel(S2b.cpu().numpy(), S2a.cpu().numpy()) < 1e-12
    report["eom_sigma_54e_65pw_vs_oracle"] = worst
    del single, sharded, cc, dVd, dVd_in

    # ---- 4. vector- / system-parallel drivers -----------------------------------------------
    g = np.load(os.path.join(ROOT, "tests", "golden", "mol_LiH_321g.npz"))
    no = int(g["n_elec"]) // 2
    dV = part_2_body_int(no, bk.asdev(g["V"].copy()))
    cc = ccsd.CCSD(no)
    T1d, T2d = bk.asdev(g["ccsd_t1"].copy()), bk.asdev(g["ccsd_t2"].copy())
    fd = bk.asdev(g["fock"].copy())
    ft = cc.get_T1_dressed_fock(fd, T1d, dV)
    dVt = cc.get_T1_dressed_V(T1d, dV)
    eom = eom_ccsd.EOM_CCSD(no, n_excit=len(g["eom_e"]), comm=comm, parallel="vectors")
    roots = eom.solve(ft, dVt, T2d)
    assert np.abs(np.sort(roots) - np.sort(g["eom_e"])).max() < 1e-8, roots
    report["davidson_vectors_roots_err"] = float(np.abs(np.sort(roots) - np.sort(g["eom_e"])).max())
    gf = np.load(os.path.join(ROOT, "tests", "golden", "feast_LiH.npz"))

    def feast(c):
        np.random.seed(5)
        fe = feast_eom_ccsd.FEAST_EOM_CCSD(no, e_c=float(gf["e_c"]), e_r=float(gf["e_r"]), n_trial=4, max_iter=2,
                                           comm=c, parallel="systems")
        fe.n_nodes = 4
        fe.max_systems = 3
        return np.sort_complex(np.asarray(fe.solve(ft, dVt, T2d)))
    ev_one, ev_dealt = feast(None), feast(comm)
    assert np.abs(ev_one - ev_dealt).max() < 1e-7, (ev_one, ev_dealt)
    report["feast_systems_vs_single_err"] = float(np.abs(ev_one - ev_dealt).max())
    report["launches"] = bk.launch_count() - launches0
    report["nccl_version"] = ".".join(str(x) for x in torch.cuda.nccl.version())
    if rank == 0:
        print("NCCL_PARITY_OK " + json.dumps(report), flush=True)
        if len(sys.argv) > 1:
            with open(sys.argv[1], "w") as fh:
                json.dump(report, fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
