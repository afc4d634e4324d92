#!/bin/bash
# Round 2, call 10 (one GPU): C4 with the diagonal preconditioner (extension) -- does Davidson converge?
mkdir -p gpurun_out
timeout 500 python tools/bench_excited.py davidson 13 10 80 diagonal > gpurun_out/r2_c4_davidson_54e_203_diag.json 2> gpurun_out/r2_c4_203_diag.log
tail -2 gpurun_out/r2_c4_203_diag.log; cat gpurun_out/r2_c4_davidson_54e_203_diag.json | cut -c1-1200
timeout 700 python tools/bench_excited.py davidson 20 10 30 diagonal > gpurun_out/r2_c4_davidson_54e_389_diag.json 2> gpurun_out/r2_c4_389_diag.log
tail -2 gpurun_out/r2_c4_389_diag.log; cat gpurun_out/r2_c4_davidson_54e_389_diag.json | cut -c1-1200
