#!/bin/bash
# Round 2, fifth GPU call (one GPU): validation of the round's kernels + the evidence for profiles/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_e.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_e.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.log
tail -2 gpurun_out/r2_bench_n1_e.log; cat gpurun_out/r2_bench_n1_e.json | cut -c1-260
timeout 600 python tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n1_e > /dev/null 2>&1
head -24 gpurun_out/r2_sweep_profile_n1_e_rank0.txt | cut -c1-130; grep -A16 "kernels by name" gpurun_out/r2_sweep_profile_n1_e_rank0.txt | cut -c1-140
timeout 300 python tools/bench_hbm.py > gpurun_out/r2_hbm.txt 2>&1; cp gpurun_out/hbm_kernels.json gpurun_out/r2_hbm_kernels.json; grep -o "^[a-z_0-9A-Z]* \|'GBps': [0-9.]*\|'frac_of_measured_peak': [0-9.]*" gpurun_out/r2_hbm.txt | paste - - - | head -20
HBM_REPS=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_hbm_ncu.csv python tools/bench_hbm.py > gpurun_out/r2_hbm_ncu.log 2>&1
wc -l gpurun_out/r2_hbm_ncu.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-calibration > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_bench_n1.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:contract_ws -c 1 -f -o gpurun_out/r2_pp_gen_v488 python tools/profile_pp_virtual.py 25 1 > gpurun_out/r2_ncu_pp.log 2>&1
tail -2 gpurun_out/r2_ncu_pp.log; ls -la gpurun_out/r2_pp_gen_v488.ncu-rep
