#!/bin/bash
# Round 2, call 14 (one GPU): ncu --set full capture of one ring-type launch of contract_ws_kernel (the dominant kernel now).
mkdir -p gpurun_out
timeout 300 python tools/profile_ring.py 488 3 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:contract_ws -c 1 -f -o gpurun_out/r2_ring_ws python tools/profile_ring.py 488 1 > gpurun_out/r2_ncu_ring.log 2>&1
tail -2 gpurun_out/r2_ncu_ring.log; ls -la gpurun_out/r2_ring_ws.ncu-rep
