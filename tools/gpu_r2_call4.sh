#!/bin/bash
# Round 2, fourth GPU call (one GPU): A/B of the consumer stagger and of the operand-role swap.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_d.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_d.txt
for ns in 0 250 500 1000 2000 4000; do
  echo "stagger $ns ns:" >> gpurun_out/r2_ab_stagger.txt
  PMB_WS_STAGGER_NS=$ns timeout 300 python tools/profile_pp_virtual.py 20 3 >> gpurun_out/r2_ab_stagger.txt 2>&1
  PMB_WS_STAGGER_NS=$ns timeout 300 python tools/profile_pp_virtual.py 13 5 dense >> gpurun_out/r2_ab_stagger.txt 2>&1
done
cat gpurun_out/r2_ab_stagger.txt
PYMES_B200_SMALL_SIDE_TO_N=0 timeout 300 python tools/profile_sweep.py 20 virtual gpurun_out/r2_ab_swap0 > /dev/null 2>&1
PYMES_B200_SMALL_SIDE_TO_N=1 timeout 300 python tools/profile_sweep.py 20 virtual gpurun_out/r2_ab_swap1 > /dev/null 2>&1
grep -h "sweep \|ci,abcj\|cj,iacb\|abid,dj\|iabc,cj" gpurun_out/r2_ab_swap0_rank0.txt gpurun_out/r2_ab_swap1_rank0.txt
