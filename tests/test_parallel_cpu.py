"""World-size-2 (gloo, CPU) tests of the (ab)-row-block sharding in pymes_b200/parallel.py.
The C ABI is emulated in numpy (tests/abi_emulator.py); the NCCL path on the GPUs runs the
same Python with device tensors."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _worker(rank, world, port, tag, is_dcsd, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import abi_emulator
        abi_emulator.install(_Patch())
        from pymes_b200 import log, parallel
        from pymes_b200.integral.partition import part_2_body_int
        log.set_quiet(True)
        g = np.load(os.path.join(ROOT, "tests", "golden", "mol_%s.npz" % tag))
        no = int(g["n_elec"]) // 2
        comm = parallel.Comm()
        nv = g["fock"].shape[0] - no
        shard = parallel.Shard(comm, nv)
        dV = part_2_body_int(no, torch.from_numpy(g["V"].copy()))
        cc = parallel.ShardedCCSD(no, comm, is_dcsd=is_dcsd)
        if is_dcsd:      # either the dictionary of local row blocks ...
            r = cc.solve(g["fock"], parallel.shard_blocks(dV, shard), delta_e=1e-12, max_iter=200)
        else:            # ... or the reference's dense V_pqrs, sliced per rank inside
            r = cc.solve(g["fock"], g["V"].copy(), delta_e=1e-12, max_iter=200)
        q.put((rank, float(r["ccsd e"]), r["t1"].numpy().copy(), r["t2"].numpy().copy(), cc.iterations,
               shard.lo, shard.na))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tag,is_dcsd", [("LiH_321g", False), ("LiH_tc", True), ("LiH_321g", True)])
def test_sharded_ccsd_world2_matches_reference(tag, is_dcsd):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + (7 if is_dcsd else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tag, is_dcsd, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "mol_%s.npz" % tag))
    name = "dcsd" if is_dcsd else "ccsd"
    for rank, e, t1, t2, its, lo, na in res:
        assert abs(e - g[name + "_e"]) < 1e-10
        assert its == len(g[name + "_trace"])
        np.testing.assert_allclose(t1, g[name + "_t1"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(t2, g[name + "_t2"], rtol=1e-8, atol=1e-11)
    assert sorted(r[5] for r in res) == [0, res[0][6] if res[0][5] == 0 else res[1][6]]


def _ueg_case():
    """TC-UEG 14e in 19 plane waves: Fock matrix, model and the integral kinds of the blocks."""
    from pymes_b200.model import ueg
    m = ueg.UEG(14, 7, 7, 0.5)
    m.init_single_basis(2.0)
    m.gamma, m.k_cutoff = None, 1.0
    parts = [("only_non_hermi_2b", m.trunc), ("effect_2b", m.trunc)]
    rng = np.random.default_rng(2)
    nP = m.n_orb
    fock = np.diag(m.kinetic()) + 1e-3 * rng.standard_normal((nP, nP))
    return m, parts, fock


def _ueg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import abi_emulator
        abi_emulator.install(_Patch())
        from pymes_b200 import log, parallel
        from pymes_b200.model.ueg import VirtualBlock
        log.set_quiet(True)
        m, parts, fock = _ueg_case()
        comm = parallel.Comm()
        dV = parallel.build_sharded_hamiltonian(m, 7, comm, parts, virtual=("abcd",))
        assert isinstance(dV["abcd"], VirtualBlock) and dV["abcd"].nz is not None
        cc = parallel.ShardedCCSD(7, comm)
        cc.setup(fock, dV)
        assert dV["abcd"].shape[0] == cc.shard.na
        es = [sum(cc.sweep()[:3]) for _ in range(3)]
        q.put((rank, es, cc._st["T1"].numpy().copy(), cc._st["T2"].numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_ccsd_world2_generated_abcd_rows(cpu_abi):
    """Each rank GENERATES its (ab) row block of V_abcd (pmb_ueg_operand_t, compressed values)
    instead of holding it: three CCSD sweeps equal the single-process solve on dense blocks."""
    from pymes_b200.integral.partition import KEYS
    from pymes_b200.solver import ccsd
    m, parts, fock = _ueg_case()
    ref = ccsd.CCSD(7)
    ref.setup(fock, m.eval_2b_blocks(7, list(KEYS), parts))
    want = [sum(ref.sweep()[:3]) for _ in range(3)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_ueg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, es, t1, t2 in res:
        np.testing.assert_allclose(es, want, rtol=0, atol=1e-11)
        np.testing.assert_allclose(t1, ref._st["T1"].numpy(), rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(t2, ref._st["T2"].numpy(), rtol=1e-9, atol=1e-13)


def _eom_inputs():
    """LiH/3-21G: dressed Fock / integrals from the golden CCSD amplitudes (test_eom_ccsd.py:30-47)."""
    from pymes_b200.integral.partition import part_2_body_int
    from pymes_b200.solver import ccsd
    g = np.load(os.path.join(ROOT, "tests", "golden", "mol_LiH_321g.npz"))
    no = int(g["n_elec"]) // 2
    dV = part_2_body_int(no, torch.from_numpy(g["V"].copy()))
    cc = ccsd.CCSD(no)
    T1, T2 = torch.from_numpy(g["ccsd_t1"].copy()), torch.from_numpy(g["ccsd_t2"].copy())
    fock = torch.from_numpy(g["fock"].copy())
    ft = cc.get_T1_dressed_fock(fock, T1, dV)
    dVd = cc.get_T1_dressed_V(T1, dV)
    return g, no, ft, dVd, T2


def _eom_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import abi_emulator
        abi_emulator.install(_Patch())
        from pymes_b200 import log, parallel
        from pymes_b200.solver import eom_ccsd
        log.set_quiet(True)
        g, no, ft, dVd, T2 = _eom_inputs()
        comm = parallel.Comm()
        eom = eom_ccsd.EOM_CCSD(no, n_excit=len(g["eom_e"]), comm=comm)
        rng = np.random.default_rng(9)
        nv = T2.shape[0]
        U1 = torch.from_numpy(rng.standard_normal((3, nv, no)))
        U2 = torch.from_numpy(rng.standard_normal((3, nv, nv, no, no)))
        S1, S2 = eom.sigma_batched(ft, dVd, U1, U2, T2)
        assert eom._plan.shard is not None and eom._plan.shard.na < nv
        roots = eom.solve(ft, dVd, T2)
        q.put((rank, S1.numpy().copy(), S2.numpy().copy(), np.asarray(roots).copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_eom_sigma_and_davidson_world2(cpu_abi):
    """Batched EOM-CCSD sigma evaluated in (ab) row blocks over two ranks (all-gather of Ex and
    of the sigma rows) == the single-process sigma; the Davidson roots are the golden ones."""
    from pymes_b200.solver import eom_ccsd
    g, no, ft, dVd, T2 = _eom_inputs()
    rng = np.random.default_rng(9)
    nv = T2.shape[0]
    U1 = torch.from_numpy(rng.standard_normal((3, nv, no)))
    U2 = torch.from_numpy(rng.standard_normal((3, nv, nv, no, no)))
    S1, S2 = eom_ccsd.EOM_CCSD(no, n_excit=3).sigma_batched(ft, dVd, U1, U2, T2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 90)
    procs = [ctx.Process(target=_eom_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, s1, s2, roots in res:
        np.testing.assert_allclose(s1, S1.numpy(), rtol=0, atol=1e-12)
        np.testing.assert_allclose(s2, S2.numpy(), rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.sort(roots), np.sort(g["eom_e"]), rtol=0, atol=1e-8)


def _feast_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import abi_emulator
        abi_emulator.install(_Patch())
        from pymes_b200 import log, parallel
        from pymes_b200.solver import feast_eom_ccsd
        log.set_quiet(True)
        g, no, ft, dVd, T2 = _eom_inputs()
        gf = np.load(os.path.join(ROOT, "tests", "golden", "feast_LiH.npz"))
        fe = feast_eom_ccsd.FEAST_EOM_CCSD(no, e_c=float(gf["e_c"]), e_r=float(gf["e_r"]), n_trial=2, max_iter=3,
                                           comm=parallel.Comm())
        d1, d2 = fe.get_diag_singles(ft, dVd, T2), fe.get_diag_doubles(ft, dVd, T2)
        fe.u_singles, fe.u_doubles = [gf["u1"].copy()], [gf["u2"].copy()]
        q1, q2 = fe._gcrotmk(0, complex(gf["z"]), d1, d2, ft, dVd, T2)
        assert fe._plan.shard is not None
        q.put((rank, np.asarray(q1).copy(), np.asarray(q2).copy(), float(fe.ls_residuals[0])))
    finally:
        dist.destroy_process_group()


def test_sharded_feast_linear_solve_world2(cpu_abi):
    """One shifted FEAST linear solve with the sigma product sharded over two ranks reproduces
    the reference's GCROT iterate (feast_eom_ccsd.py:293-350) like the single-process solve."""
    gf = np.load(os.path.join(ROOT, "tests", "golden", "feast_LiH.npz"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + (os.getpid() % 90)
    procs = [ctx.Process(target=_feast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    nrm = np.sqrt(np.sum(abs(gf["q1"]) ** 2) + np.sum(abs(gf["q2"]) ** 2))
    for rank, q1, q2, resid in res:
        assert resid < 1e-4
        err = np.sqrt(np.sum(abs(q1 - gf["q1"]) ** 2) + np.sum(abs(q2 - gf["q2"]) ** 2)) / nrm
        assert err < 1e-10, err


# --------------------------------------------------------------------------
# the "replicated operator, work dealt out" modes (C4 / C5): parallel="vectors" / "systems"
# --------------------------------------------------------------------------
def _feast_seeded(no, ft, dVd, T2, comm=None):
    from pymes_b200.solver import feast_eom_ccsd
    gf = np.load(os.path.join(ROOT, "tests", "golden", "feast_LiH.npz"))
    np.random.seed(5)
    fe = feast_eom_ccsd.FEAST_EOM_CCSD(no, e_c=float(gf["e_c"]), e_r=float(gf["e_r"]), n_trial=4, max_iter=2,
                                       comm=comm, parallel="systems")
    fe.n_nodes = 4
    fe.max_systems = 3              # groups of 3 systems: exercises the grouping as well
    ev = fe.solve(ft, dVd, T2)
    return np.sort_complex(np.asarray(ev)), fe


def _dealt_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import abi_emulator
        abi_emulator.install(_Patch())
        from pymes_b200 import log, parallel
        from pymes_b200.solver import eom_ccsd
        log.set_quiet(True)
        g, no, ft, dVd, T2 = _eom_inputs()
        comm = parallel.Comm()
        eom = eom_ccsd.EOM_CCSD(no, n_excit=len(g["eom_e"]), comm=comm, parallel="vectors")
        eom.max_rhs = 1
        roots = eom.solve(ft, dVd, T2)
        assert eom._plan.shard is None and eom.vec_comm is not None
        if rank == 1:
            np.random.seed(12345)   # the start vectors must come from rank 0 alone
        ev, fe = _feast_seeded(no, ft, dVd, T2, comm)
        assert fe.sys_comm is not None and fe._plan.shard is None
        q.put((rank, np.asarray(roots).copy(), ev, [t["systems_this_rank"] for t in fe.timings]))
    finally:
        dist.destroy_process_group()


def test_vector_and_system_parallel_world2(cpu_abi):
    """Davidson with the new trial vectors dealt out over two ranks gives the golden roots; a
    seeded FEAST run with the (node x trial vector) systems dealt out over two ranks gives the
    eigenvalues of the single-process run (same start vectors, sums reordered only)."""
    g, no, ft, dVd, T2 = _eom_inputs()
    ev1, fe1 = _feast_seeded(no, ft, dVd, T2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 90)
    procs = [ctx.Process(target=_dealt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, roots, ev, nsys in res:
        np.testing.assert_allclose(np.sort(roots), np.sort(g["eom_e"]), rtol=0, atol=1e-8)
        # the projected generalised eigenproblem (non-orthogonal Q) amplifies the reordering round-off
        np.testing.assert_allclose(ev, ev1, rtol=0, atol=1e-7)
        assert nsys == [t["systems_this_rank"] // 2 for t in fe1.timings]
