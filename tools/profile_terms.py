"""Time every contraction launch of the CCD doubles residual separately (random dense
operands, CUDA events) and print TFLOP/s per launch.  Not a bench value: it shows which
index patterns the gather pipeline handles badly.
usage: profile_terms.py [v] [o] [reps] [only_tag|-] [panel_MB] [cfg]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymes_b200 import backend as bk

v = int(sys.argv[1]) if len(sys.argv) > 1 else 314
o = int(sys.argv[2]) if len(sys.argv) > 2 else 27
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
only = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
if len(sys.argv) > 5:
    from pymes_b200 import _lib
    _lib.load().pmb_contract_set_panel_bytes(int(float(sys.argv[5]) * (1 << 20)))
if len(sys.argv) > 6:
    from pymes_b200 import _lib
    _lib.load().pmb_contract_set_tuning(int(sys.argv[6]), 0)
torch.cuda.set_device(0)
torch.manual_seed(0)


def rnd(*s):
    return torch.randn(*s, dtype=torch.float64, device="cuda")


T2, Tt = rnd(v, v, o, o), rnd(v, v, o, o)
V_ijab, V_iajb, V_iabj = rnd(o, o, v, v), rnd(o, v, o, v), rnd(o, v, v, o)
V_klij, I = rnd(o, o, o, o), rnd(o, o, o, o)
X1, Xai, Xp = rnd(v, o, v, o), rnd(v, v, o, o), rnd(v, o, v, o)
Xac, Xki = rnd(v, v), rnd(o, o)
R = rnd(v, v, o, o)

CASES = [
    ("I_klij", "klij", [(1.0, "klcd", V_ijab, "cdij", T2)], 2 * o**4 * v**2),
    ("hh", "abij", [(1.0, "abkl", T2, "klij", I)], 2 * o**4 * v**2),
    ("X1", "alcj", [(1.0, "klcd", V_ijab, "adkj", T2)], 2 * o**3 * v**3),
    ("X1.T", "abij", [(1.0, "alcj", X1, "cbil", T2)], 2 * o**3 * v**3),
    ("Xai", "cbkj", [(1.0, "klcd", V_ijab, "dblj", Tt)], 2 * o**3 * v**3),
    ("Tt.Xai", "abij", [(1.0, "acik", Tt, "cbkj", Xai)], 2 * o**3 * v**3),
    ("Xac", "ac", [(-1.0, "adkl", Tt, "lkdc", V_ijab)], 2 * o**2 * v**3),
    ("Xki", "ki", [(1.0, "cdil", Tt, "lkdc", V_ijab)], 2 * o**3 * v**2),
    ("Xac.T", "abij", [(1.0, "ac", Xac, "cbij", T2)], 2 * o**2 * v**3),
    ("Xki.T", "abij", [(-1.0, "ki", Xki, "abkj", T2)], 2 * o**3 * v**2),
    ("Xp", "alci", [(1.0, "klcd", V_ijab, "daki", T2)], 2 * o**3 * v**3),
    ("ring4", "abij", [(-1.0, "kaic", V_iajb, "cbkj", T2), (1.0, "acik", Tt, "kbcj", V_iabj),
                       (-1.0, "alci", Xp, "cblj", T2), (1.0, "alci", Xp, "bclj", T2)], 8 * o**3 * v**3),
    ("ring234", "abij", [(-1.0, "kbic", V_iajb, "ackj", T2)], 2 * o**3 * v**3),
]
for tag, out_sub, terms, flops in CASES:
    if only and tag != only:
        continue
    shape = {"klij": (o, o, o, o), "abij": (v, v, o, o), "alcj": (v, o, v, o), "cbkj": (v, v, o, o),
             "ac": (v, v), "ki": (o, o), "alci": (v, o, v, o)}[out_sub]
    out = torch.zeros(*shape, dtype=torch.float64, device="cuda")
    before = bk.launch_count()
    bk.contract_terms(out_sub, terms, out=out, beta=1.0)
    nl = bk.launch_count() - before
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        bk.contract_terms(out_sub, terms, out=out, beta=1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-8s %-5s launches=%3d  %9.3f ms  %6.2f TFLOP/s" % (tag, out_sub, nl, ms, flops / ms / 1e9), flush=True)
