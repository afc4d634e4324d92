#!/bin/bash
# First GPU call of round 2 (one GPU): everything that was written after round 1's GPU budget was spent.
#   1. the not-yet-validated CUDA cases (tests/test_gpu_next.py) and the full -m gpu suite
#   2. EOM sigma with a never-materialised V_abcd at 389 and 515 orbitals (C4 sizes)
#   3. the DCSD variant of the bench workload
# Then, on 2 GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#                       --master-port 29520 tools/bench_eom_sharded.py 20 2 8
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_next.py -q -m gpu_next ) > gpurun_out/pytest_gpu_next.txt 2>&1
tail -5 gpurun_out/pytest_gpu_next.txt
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt
timeout 900 python tools/bench_eom.py 20 2 8 virtual > gpurun_out/eom_sigma_54e_389_virtual.json 2> gpurun_out/eom_389.log
tail -5 gpurun_out/eom_389.log
timeout 1500 python tools/bench_eom.py 25 2 4 virtual > gpurun_out/eom_sigma_54e_515_virtual.json 2> gpurun_out/eom_515.log
tail -5 gpurun_out/eom_515.log
timeout 900 python bench.py --dcsd --no-cpu > gpurun_out/bench_n1_dcsd.json 2> gpurun_out/bench_n1_dcsd.log
cat gpurun_out/bench_n1_dcsd.json
