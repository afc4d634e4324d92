#!/bin/bash
# Round 2, call 13 (one GPU): stored V blocks with geometry tags -- ring products with V_ijab momentum-blocked.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "momentum_blocked or lockstep" ) > gpurun_out/r2_pytest_gpu_blocked2.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_blocked2.txt
if ! tail -1 gpurun_out/r2_pytest_gpu_blocked2.txt | grep -q passed || tail -1 gpurun_out/r2_pytest_gpu_blocked2.txt | grep -q failed; then
  echo "blocked tests failed: stopping"; grep -n "Error\|assert" gpurun_out/r2_pytest_gpu_blocked2.txt | head -20; exit 1
fi
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_l.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_l.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_n1_blocked2.json 2> gpurun_out/r2_bench_n1_blocked2.log
tail -2 gpurun_out/r2_bench_n1_blocked2.log; cut -c1-400 gpurun_out/r2_bench_n1_blocked2.json
timeout 600 python tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n1_blocked2 > /dev/null 2>&1
head -40 gpurun_out/r2_sweep_profile_n1_blocked2_rank0.txt | cut -c1-130
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2_launches_bench_n1_blocked2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_bench_blocked2.log 2>&1
wc -l gpurun_out/r2_launches_bench_n1_blocked2.csv
