#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contract" > gpurun_out/pytest_contract.txt 2>&1
tail -5 gpurun_out/pytest_contract.txt
for cfg in 5 6 0 2; do
  timeout 200 python tools/profile_pp.py 314 27 3 $cfg >> gpurun_out/pp_ws.txt 2>&1
done
timeout 200 python tools/profile_pp.py 200 27 3 5 >> gpurun_out/pp_ws.txt 2>&1
timeout 200 python tools/profile_pp.py 314 27 3 5 0 >> gpurun_out/pp_ws.txt 2>&1
cat gpurun_out/pp_ws.txt
timeout 300 python tools/profile_terms.py 314 27 2 > gpurun_out/terms_314_ws.txt 2>&1
cat gpurun_out/terms_314_ws.txt
