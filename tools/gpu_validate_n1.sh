#!/bin/bash
# round-1 call 11: validation of the compressed generated V_abcd path + evidence for profiles/
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
tail -4 gpurun_out/bench_n1.log; cat gpurun_out/bench_n1.json
timeout 600 python tools/profile_sweep.py 25 > gpurun_out/sweep_profile_515.txt 2>&1
head -30 gpurun_out/sweep_profile_515.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_bench_n1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_bench_n1.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:contract_ws -c 1 -f -o gpurun_out/pp_gen_v488 python tools/profile_pp_virtual.py 25 1 > gpurun_out/ncu_pp_gen488.log 2>&1
tail -2 gpurun_out/ncu_pp_gen488.log
timeout 300 python tools/bench_hbm.py > gpurun_out/hbm.txt 2>&1; tail -3 gpurun_out/hbm.txt
