#!/bin/bash
# 2-GPU: block-cyclic dense path timings (NCCL), C3-sized problem
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 tests/sharded_bench.py --size 60000 --grad --check 2> gpurun_out/sharded_60k.err | tail -1 | tee gpurun_out/sharded_60k.json; grep -E "Error|error" gpurun_out/sharded_60k.err | head -5
timeout 900 $TR --master-port 29552 tests/sharded_bench.py --size 100000 --grad 2> gpurun_out/sharded_100k.err | tail -1 | tee gpurun_out/sharded_100k.json; grep -E "Error|error" gpurun_out/sharded_100k.err | head -5
timeout 600 $TR --master-port 29553 tests/sharded_bench.py --size 60000 --nb 1024 2> gpurun_out/sharded_60k_nb1024.err | tail -1 | tee gpurun_out/sharded_60k_nb1024.json
timeout 300 $TR --master-port 29554 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2> gpurun_out/bench_ref_2gpu.err | tail -1 | cut -c1-200
