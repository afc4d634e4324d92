#!/bin/bash
# round-1 call 12: m-fast tile order for the generated ladder (DRAM traffic), pmb_gemv, EOM sigma batches
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python tools/profile_pp_virtual.py 25 2 > gpurun_out/pp_virtual.txt 2>&1
cat gpurun_out/pp_virtual.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:contract_ws -c 1 --csv --log-file gpurun_out/pp_gen_v488_traffic.csv python tools/profile_pp_virtual.py 25 1 > gpurun_out/ncu_pp_traffic.log 2>&1
tail -8 gpurun_out/pp_gen_v488_traffic.csv | cut -c1-400
timeout 600 python tools/profile_sweep.py 25 > gpurun_out/sweep_profile_515.txt 2>&1
head -24 gpurun_out/sweep_profile_515.txt
timeout 900 python tools/bench_eom.py 13 3 64 > gpurun_out/eom_sigma_54e_203.json 2> gpurun_out/eom_sigma_54e_203.log
tail -8 gpurun_out/eom_sigma_54e_203.log
