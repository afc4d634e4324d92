"""Per-contraction timing of ONE CCSD+DIIS sweep on the bench workload (TC-UEG 54e, N=1):
every pmb_contract launch bracketed by CUDA events, grouped by index pattern.  Not a bench
value (the extra events serialise nothing, but the run is a diagnostic, not the timed step).
usage: profile_sweep.py [cutoff] [dense]   (dense: store V_abcd instead of generating it)"""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pymes_b200 import backend as bk, log as plog
from pymes_b200.model import ueg
from pymes_b200.solver import ccsd
from pymes_b200.integral.partition import KEYS

cutoff = float(sys.argv[1]) if len(sys.argv) > 1 else bench.CUTOFF_FOR_GPUS[1]
torch.cuda.set_device(0)
plog.set_quiet(True)
no = bench.N_ELE // 2
m = ueg.UEG(bench.N_ELE, no, no, bench.RS)
m.init_single_basis(cutoff)
m.k_cutoff, m.gamma = bench.K_CUTOFF, None
fock = bench.build_fock(m, no)
cc = ccsd.CCSD(no)
dense = len(sys.argv) > 2 and sys.argv[2] == "dense"
dV = m.eval_2b_blocks(no, list(KEYS), bench.tc_parts(m), virtual=() if dense else ("abcd",))
cc.setup(fock, dV)
for _ in range(2):
    cc.sweep()
torch.cuda.synchronize()
bk.enable_trace(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
cc.sweep()
ev1.record()
rows = bk.trace_report()
bk.enable_trace(False)
total = ev0.elapsed_time(ev1)
agg = collections.OrderedDict()
for lab, fl, ms in rows:
    a = agg.setdefault(lab, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += fl
    a[2] += ms
print("nP=%d o=%d v=%d: sweep %.1f ms, %d contraction calls, %.1f ms inside them"
      % (m.n_orb, no, m.n_orb - no, total, len(rows), sum(r[2] for r in rows)))
for lab, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("%9.3f ms  x%-2d %7.2f TFLOP/s  %s" % (ms, n, fl / ms / 1e9 if ms > 0 else 0.0, lab))
