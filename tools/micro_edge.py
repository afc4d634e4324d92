"""Does the ragged last column of tiles cost its padded or its real width?  C[M,N] = A[M,K] B[K,N]
on the warp-specialised kernel for N = 640 (5 full tiles), 729 (the pp ladder: 5 + 89 columns),
768 (6 full tiles); same M, K.  usage: micro_edge.py [v=200]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymes_b200 import backend as bk, _lib

v = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.cuda.set_device(0)
M = K = v * v
A = torch.randn(M, K, dtype=torch.float64, device="cuda")
_lib.load().pmb_contract_set_tuning(5, 0)
for N in (640, 704, 729, 736, 768):
    B = torch.randn(K, N, dtype=torch.float64, device="cuda")
    C = torch.empty(M, N, dtype=torch.float64, device="cuda")
    bk.contract("mk,kn->mn", A, B, out=C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bk.contract("mk,kn->mn", A, B, out=C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("N=%4d  %8.3f ms  %6.2f TFLOP/s (real N)  %6.2f TFLOP/s (padded to %d)"
          % (N, ms, 2.0 * M * K * N / ms / 1e9, 2.0 * M * K * (-(-N // 128) * 128) / ms / 1e9, -(-N // 128) * 128))
