"""The particle-particle ladder with V_abcd never materialised (pmb_ueg_operand_t), alone:
TC-UEG 54e pair tables, random tau.  For `ncu -k regex:contract_ws_kernel` and for the
generated-vs-stored comparison.  usage: profile_pp_virtual.py [cutoff] [reps] [dense]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pymes_b200 import backend as bk, log as plog
from pymes_b200.model import ueg

cutoff = float(sys.argv[1]) if len(sys.argv) > 1 else 25.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
torch.cuda.set_device(0)
plog.set_quiet(True)
no = bench.N_ELE // 2
m = ueg.UEG(bench.N_ELE, no, no, bench.RS)
m.init_single_basis(cutoff)
m.k_cutoff, m.gamma = bench.K_CUTOFF, None
nv = m.n_orb - no
V = m.eval_2b_blocks(no, ["abcd"], bench.tc_parts(m), virtual=() if dense else ("abcd",))["abcd"]
g = torch.Generator(device="cuda").manual_seed(0)
T = torch.randn(nv, nv, no, no, dtype=torch.float64, device="cuda", generator=g)
R = torch.zeros(nv, nv, no, no, dtype=torch.float64, device="cuda")
bk.contract("abcd,cdij->abij", V, T, out=R, beta=1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    bk.contract("abcd,cdij->abij", V, T, out=R, beta=1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("pp ladder nP=%d v=%d o=%d V_abcd %s: %.3f ms, %.2f TFLOP/s"
      % (m.n_orb, nv, no, "stored" if dense else "generated", ms, 2.0 * nv ** 4 * no * no / ms / 1e9))
