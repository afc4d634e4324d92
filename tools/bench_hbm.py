"""Achieved HBM bandwidth of the HBM-bound kernels (CUDA events, inputs >> L2).
Writes gpurun_out/hbm_kernels.json.  ALGORITHMIC bytes per element:
update 32 B (read R, read T, write T, write dT), energy 24 B (the row of T2 and the rows (a,b), (b,a)
of the [a,b,i,j]-stored V_ijab; 16 B if the exchange rows were paired), DIIS dots 8(m+1) B, DIIS combine 8(m+1) B, UEG build 8 B per stored element,
tilde 16 B (row pairs: every row read once), sym_baji 24 B (Ex once, R read + written), strided
copy / transpose 16 B."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pymes_b200 import backend as bk
from pymes_b200.model import ueg

torch.cuda.set_device(0)
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
no, nv = 27, 314
n = nv * nv * no * no
res = {"shape": "o=27 v=314 (%.2f GB per T2-sized tensor)" % (n * 8 / 1e9), "peak_gbs": PEAK}


REPS = int(os.environ.get("HBM_REPS", "5"))       # HBM_REPS=1 under ncu


def timeit(fn, reps=None):
    reps = REPS if reps is None else reps
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def rec(name, seconds, nbytes):
    res[name] = {"ms": seconds * 1e3, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / seconds / 1e9,
                 "frac_of_measured_peak": nbytes / seconds / 1e9 / PEAK}
    print(name, res[name])


g = torch.Generator(device="cuda").manual_seed(0)
T = torch.randn(nv, nv, no, no, dtype=torch.float64, device="cuda", generator=g)
R = torch.randn_like(T)
V = torch.randn(no, no, nv, nv, dtype=torch.float64, device="cuda", generator=g)
ei = torch.linspace(-2, -1, no, dtype=torch.float64, device="cuda")
ea = torch.linspace(1, 3, nv, dtype=torch.float64, device="cuda")
scal = bk.zeros(8)
rec("update_doubles", timeit(lambda: bk.update_doubles(ei, ea, 0.0, 1.0, R, T, scal[3:4])), 32 * n)
from pymes_b200.solver import ccsd
Ve = ccsd.energy_layout(V)       # what CCSD.sweep hands over: V_ijab stored as [a,b,i,j]
rec("energy_doubles", timeit(lambda: bk.energy_doubles(T, Ve, scal)), 24 * n)      # T2 row + V rows (a,b) and (b,a)
rec("energy_doubles_ijab_storage", timeit(lambda: bk.energy_doubles(T, V, scal)), 16 * n)
rec("tilde", timeit(lambda: bk.tilde(T)), 16 * n)
rec("sym_baji", timeit(lambda: bk.sym_baji(T, R, accumulate=True)), 24 * n)
xs = [torch.randn_like(T) for _ in range(6)]
rec("diis_dots_m6", timeit(lambda: bk.dots(xs, R)), 8 * 7 * n)
c = list(np.linspace(0.1, 0.6, 6))
rec("diis_lincomb_m6", timeit(lambda: bk.lincomb(c, xs)), 8 * 7 * n)
rec("copy_strided_V_block", timeit(lambda: bk.copy(V.permute(2, 3, 0, 1))), 16 * n)
del xs
m = ueg.UEG(54, 27, 27, 1.0)
m.init_single_basis(10.0)
m.k_cutoff, m.gamma = 2.0, None
nP = m.n_orb
W0, W1 = m.pair_tables("only_2b", m.trunc)
out = bk.empty(nP, nP, nP, nP)
rec("ueg_build_block_dense_nP147", timeit(lambda: m.build_block((0,) * 4, (nP,) * 4, W0a=W0, W1a=W1, out=out)),
    8 * nP ** 4)
t0 = timeit(lambda: m.pair_tables("only_2b", m.trunc), reps=1)
res["ueg_pair_tables_only_2b_nP147_ms"] = t0 * 1e3
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/hbm_kernels.json", "w"), indent=1)
