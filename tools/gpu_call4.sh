#!/bin/bash
mkdir -p gpurun_out
python -c "from pymes_b200 import build as b; print(b.build_library())"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python tools/profile_pp.py 314 27 3 5 > gpurun_out/pp_ws.txt 2>&1
timeout 200 python tools/profile_pp.py 200 27 3 5 >> gpurun_out/pp_ws.txt 2>&1
cat gpurun_out/pp_ws.txt
timeout 300 python tools/profile_terms.py 314 27 2 > gpurun_out/terms_314_ws3.txt 2>&1
cat gpurun_out/terms_314_ws3.txt
timeout 600 python bench.py > gpurun_out/bench_r1_n1_ws.json 2> gpurun_out/bench_r1_n1_ws.log
cat gpurun_out/bench_r1_n1_ws.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_ws.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
