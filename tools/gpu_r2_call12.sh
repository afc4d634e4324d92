#!/bin/bash
# Round 2, call 12 (one GPU): the momentum-blocked ladder (pmb_blocked_contract) -- parity, bench A/B, profiles.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "momentum_blocked or virtual or o27 or lockstep" ) > gpurun_out/r2_pytest_gpu_blocked.txt 2>&1
tail -5 gpurun_out/r2_pytest_gpu_blocked.txt
if ! tail -1 gpurun_out/r2_pytest_gpu_blocked.txt | grep -q passed || tail -1 gpurun_out/r2_pytest_gpu_blocked.txt | grep -q failed; then
  echo "blocked tests failed: stopping"; grep -n "Error\|assert" gpurun_out/r2_pytest_gpu_blocked.txt | head -20; exit 1
fi
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_k.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_k.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_n1_blocked.json 2> gpurun_out/r2_bench_n1_blocked.log
tail -2 gpurun_out/r2_bench_n1_blocked.log; cut -c1-400 gpurun_out/r2_bench_n1_blocked.json
timeout 600 python bench.py --ladder dense --steps 2 --warmup 2 --no-cpu --no-calibration > gpurun_out/r2_bench_n1_dense_ab.json 2> gpurun_out/r2_bench_n1_dense_ab.log
cut -c1-300 gpurun_out/r2_bench_n1_dense_ab.json
timeout 600 python tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n1_blocked > /dev/null 2>&1
head -30 gpurun_out/r2_sweep_profile_n1_blocked_rank0.txt | cut -c1-130
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2_launches_bench_n1_blocked.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_bench_blocked.log 2>&1
wc -l gpurun_out/r2_launches_bench_n1_blocked.csv
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:blocked_kernel -c 3 -f -o gpurun_out/r2_blocked_ladder python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_blocked.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:contract_ws -c 2 -f -o gpurun_out/r2_ring_ws python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_ring.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
