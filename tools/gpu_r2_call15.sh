#!/bin/bash
# Round 2, call 15 (one GPU): banded tile order of the warp-specialised kernel (stored operands) -- parity, A/B, traffic.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_m.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_m.txt
if ! tail -1 gpurun_out/r2_pytest_gpu_m.txt | grep -q passed || tail -1 gpurun_out/r2_pytest_gpu_m.txt | grep -q failed; then
  echo "tests failed: stopping"; grep -n "Error\|assert" gpurun_out/r2_pytest_gpu_m.txt | head -20; exit 1
fi
timeout 300 python tools/profile_ring.py 488 3 2>&1 | tail -1
PYMES_B200_TUNING=261 timeout 300 python tools/profile_ring.py 488 3 2>&1 | tail -1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off -k regex:contract_ws -c 1 --csv --log-file gpurun_out/r2_ring_banded_ncu.csv python tools/profile_ring.py 488 1 > /dev/null 2>&1
tail -2 gpurun_out/r2_ring_banded_ncu.csv | cut -c1-400
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench_n1_banded.json 2> gpurun_out/r2_bench_n1_banded.log
cut -c1-300 gpurun_out/r2_bench_n1_banded.json
