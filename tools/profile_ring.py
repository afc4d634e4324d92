"""One ring-type contraction of the doubles residual (ccd.py:234, "kbic,ackj->abij": 2 o^3 v^3 flop) at the
bench shape (o = 27, v = 488) on random operands, for an `ncu --set full -k regex:contract_ws -c 1` capture
of the kernel that dominates the iteration once the integral-block products run momentum-blocked.
    python tools/profile_ring.py [v] [reps]
Diagnostic: prints the CUDA-event time per launch; a number taken under a profiler is not a bench value."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymes_b200 import backend as bk

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 488
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
no = 27
if os.environ.get("PYMES_B200_TUNING"):        # e.g. 261 = configuration 5 + 256: n-fastest tile order (A/B)
    from pymes_b200 import _lib
    _lib.load().pmb_contract_set_tuning(int(os.environ["PYMES_B200_TUNING"]), 0)
g = torch.Generator(device="cuda").manual_seed(0)
V = torch.randn(no, nv, no, nv, dtype=torch.float64, device="cuda", generator=g)
T = torch.randn(nv, nv, no, no, dtype=torch.float64, device="cuda", generator=g)
out = bk.empty(nv, nv, no, no)
bk.contract("kbic,ackj->abij", V, T, out=out)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
ev0.record()
for _ in range(reps):
    bk.contract("kbic,ackj->abij", V, T, out=out)
ev1.record()
torch.cuda.profiler.stop()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
print("kbic,ackj->abij  o=%d v=%d: %.3f ms per launch = %.2f TFLOP/s" % (no, nv, ms, 2.0 * no ** 3 * nv ** 3 / ms / 1e9))
