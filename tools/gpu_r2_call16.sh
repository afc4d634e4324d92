#!/bin/bash
# Round 2, call 16 (one GPU): final verification of the tree -- whole -m gpu suite, smoke, the default bench line,
# its ncu launch list, and the batched EOM sigma at 389 plane waves with the momentum-blocked ladder.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_n.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_n.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final2.json 2> gpurun_out/r2_bench_n1_final2.log
tail -1 gpurun_out/r2_bench_n1_final2.log; cut -c1-300 gpurun_out/r2_bench_n1_final2.json
timeout 400 python tools/bench_eom.py 20 2 4 virtual > gpurun_out/r2_eom_sigma_54e_389_blocked.json 2> gpurun_out/r2_eom_389_blocked.log
tail -3 gpurun_out/r2_eom_389_blocked.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2_launches_bench_n1_final2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_bench_final2.log 2>&1
wc -l gpurun_out/r2_launches_bench_n1_final2.csv
