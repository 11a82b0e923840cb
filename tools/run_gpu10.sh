#!/bin/bash
# 2-GPU validation: block-cyclic dense path over NCCL (tests + timings), replica bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/host2.txt; nvidia-smi topo -m 2>/dev/null | head -8 | tee -a gpurun_out/host2.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -15 | tee gpurun_out/pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 tests/sharded_bench.py --size 60000 --grad --check 2> gpurun_out/sharded_60k.err | tail -1 | tee gpurun_out/sharded_60k.json; tail -3 gpurun_out/sharded_60k.err
timeout 900 $TR --master-port 29552 tests/sharded_bench.py --size 100000 --grad 2> gpurun_out/sharded_100k.err | tail -1 | tee gpurun_out/sharded_100k.json; tail -3 gpurun_out/sharded_100k.err
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --steps 1 --warmup 3 2> gpurun_out/bench_2gpu.err | tail -1 | tee gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 300 $TR --master-port 29554 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2> gpurun_out/bench_ref_2gpu.err | tail -1 | tee gpurun_out/bench_ref_2gpu.json
