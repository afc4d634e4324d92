#!/bin/bash
mkdir -p gpurun_out
python -c "from pymes_b200 import build as b; print(b.build_library())"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contract" > gpurun_out/pytest_contract.txt 2>&1
tail -3 gpurun_out/pytest_contract.txt
rm -f gpurun_out/pp_ws.txt
for mb in 40 0 80 160 320; do
  timeout 200 python tools/profile_pp.py 314 27 3 5 $mb >> gpurun_out/pp_ws.txt 2>&1
done
cat gpurun_out/pp_ws.txt
for mb in 0 40; do
echo "panel $mb" >> gpurun_out/terms_314_ws.txt
timeout 300 python tools/profile_terms.py 314 27 2 - $mb >> gpurun_out/terms_314_ws2.txt 2>&1
done
cat gpurun_out/terms_314_ws2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_ws -c 1 -s 1 -o gpurun_out/pp_ws_v200 -f python tools/profile_pp.py 200 27 1 5 0 > gpurun_out/ncu_pp_ws.log 2>&1
tail -3 gpurun_out/ncu_pp_ws.log
