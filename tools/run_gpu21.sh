#!/bin/bash
# 2-GPU box: the whole GPU suite (incl. the 2-rank NCCL test and the new fused user-kernel gradient test) + sharded timings after the TRTRI hoist
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29552 tests/sharded_bench.py --size 100000 --grad 2> gpurun_out/sharded_100k.err | tail -1 | tee gpurun_out/sharded_100k.json; grep -E "Error|error" gpurun_out/sharded_100k.err | head -5
timeout 600 $TR --master-port 29551 tests/sharded_bench.py --size 60000 --grad --check 2> gpurun_out/sharded_60k.err | tail -1 | tee gpurun_out/sharded_60k.json; grep -E "Error|error" gpurun_out/sharded_60k.err | head -5
