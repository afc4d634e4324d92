"""Measure what the box gives for FP64: cuBLAS DGEMM (torch.matmul) and a device copy.
Run under gpurun; writes gpurun_out/fp64_calibration.json.  Not part of the product path."""
import json
import os
import torch

torch.cuda.set_device(0)
out = {}
for n in (4096, 8192, 12288):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_%d_tflops" % n] = 2 * n ** 3 / best / 1e9
# tall-skinny like the pp ladder: M=K=v^2, N=o^2
v, o = 200, 27
a = torch.randn(v * v, v * v, dtype=torch.float64, device="cuda")
b = torch.randn(v * v, o * o, dtype=torch.float64, device="cuda")
for _ in range(2):
    a @ b
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    a @ b
e1.record()
torch.cuda.synchronize()
out["dgemm_pp_v200_o27_tflops"] = 3 * 2 * (v * v) ** 2 * o * o / e0.elapsed_time(e1) / 1e9
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
y.copy_(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y.copy_(x)
e1.record()
torch.cuda.synchronize()
out["copy_gbs"] = 2 * x.numel() * 8 / e0.elapsed_time(e1) / 1e6
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_calibration.json", "w"), indent=1)
print(out)
