"""Launch the particle-particle ladder contraction R[ab,ij] += V[ab,cd] T[cd,ij] alone
(for `ncu -k regex:contract_kernel`).  usage: profile_pp.py [v] [o] [reps] [cfg]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymes_b200 import backend as bk, _lib

v = int(sys.argv[1]) if len(sys.argv) > 1 else 200
o = int(sys.argv[2]) if len(sys.argv) > 2 else 27
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else -1
panel_mb = float(sys.argv[5]) if len(sys.argv) > 5 else -1.0
torch.cuda.set_device(0)
_lib.load().pmb_contract_set_tuning(cfg, 0)
_lib.load().pmb_contract_set_panel_bytes(int(panel_mb * (1 << 20)))
V = torch.randn(v, v, v, v, dtype=torch.float64, device="cuda")
T = torch.randn(v, v, o, o, dtype=torch.float64, device="cuda")
R = torch.zeros(v, v, o, o, dtype=torch.float64, device="cuda")
bk.contract("abcd,cdij->abij", V, T, out=R, beta=1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    bk.contract("abcd,cdij->abij", V, T, out=R, beta=1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("pp ladder v=%d o=%d cfg=%d panel=%gMB: %.3f ms, %.2f TFLOP/s"
      % (v, o, cfg, panel_mb, ms, 2.0 * v ** 4 * o * o / ms / 1e9))
