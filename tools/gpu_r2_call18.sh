#!/bin/bash
# Round 2, call 18 (one GPU): batched EOM sigma at 389 plane waves, executed flops summed from the launch trace.
mkdir -p gpurun_out
timeout 200 python tools/bench_eom.py 20 2 4 virtual > gpurun_out/r2_eom_sigma_54e_389_blocked_b.json 2> gpurun_out/r2_eom_389_blocked_b.log
tail -3 gpurun_out/r2_eom_389_blocked_b.log
