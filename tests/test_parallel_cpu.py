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
        r = cc.solve(g["fock"], parallel.shard_blocks(dV, shard), delta_e=1e-12, max_iter=200)
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
