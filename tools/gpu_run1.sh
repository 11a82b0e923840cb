#!/bin/bash
# GPU run 1 (round 2): full -m gpu suite + the N=1 bench line with every sub-record
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_v1_gpu.txt 2>&1
nproc >> gpurun_out/r02_v1_gpu.txt; free -g | head -2 >> gpurun_out/r02_v1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02_v1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_v1_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_v1_bench.json 2> gpurun_out/r02_v1_bench.err
echo "bench rc=$?" >> gpurun_out/r02_v1_bench.err
tail -5 gpurun_out/r02_v1_pytest.log
