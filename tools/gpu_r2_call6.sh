#!/bin/bash
# A/B: where do the last 8 % of the DMMA peak go?  PMB_WS_DEBUG: 0 normal, 1 producers publish stages
# without copying, 2 B tiles only (timing only; the numbers computed are garbage).
mkdir -p gpurun_out
for dbg in 0 1 2; do
  echo "PMB_WS_DEBUG=$dbg" >> gpurun_out/r2_ab_wsdebug.txt
  PMB_WS_DEBUG=$dbg timeout 300 python tools/profile_pp_virtual.py 20 3 >> gpurun_out/r2_ab_wsdebug.txt 2>&1
  PMB_WS_DEBUG=$dbg timeout 300 python tools/profile_pp_virtual.py 13 5 dense >> gpurun_out/r2_ab_wsdebug.txt 2>&1
done
cat gpurun_out/r2_ab_wsdebug.txt
( timeout 900 python -m pytest tests -m gpu -x -q -k "correlator or remaining or tc_tables or ueg" ) > gpurun_out/r2_pytest_gpu_f.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_f.txt
