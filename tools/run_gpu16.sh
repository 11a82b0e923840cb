#!/bin/bash
# 8-GPU: C5 (N = 200 000 dense, block-cyclic), N = 120 000 with gradient, and the replica bench the driver runs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR8 --master-port 29561 bench.py --workload c5 --steps 1 --warmup 0 2> gpurun_out/bench_c5_8gpu.err | tail -1 | tee gpurun_out/bench_c5_8gpu.json; grep -E "Error|error|Traceback" gpurun_out/bench_c5_8gpu.err | head -5
timeout 400 $TR8 --master-port 29562 bench.py --workload c5 --size 120000 --grad --steps 1 --warmup 0 2> gpurun_out/bench_c5_8gpu_grad.err | tail -1 | tee gpurun_out/bench_c5_8gpu_grad.json; grep -E "Error|error|Traceback" gpurun_out/bench_c5_8gpu_grad.err | head -5
timeout 300 $TR8 --master-port 29563 bench.py --gpus 8 --steps 1 --warmup 3 2> gpurun_out/bench_8gpu.err | tail -1 | tee gpurun_out/bench_8gpu.json; grep -E "Error|error|Traceback" gpurun_out/bench_8gpu.err | head -5
