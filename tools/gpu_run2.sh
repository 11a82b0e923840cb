#!/bin/bash
# GPU run 2 (round 2, 2 GPUs): NCCL tests of the sharded dense + sparse paths, sharded timing harness, 2-rank bench
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --durations=8 > gpurun_out/r02_v2_pytest_sharded.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_v2_pytest_sharded.log
tail -4 gpurun_out/r02_v2_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    tests/sharded_bench.py --size 60000 --grad --check > gpurun_out/r02_v2_sharded_2gpu_60k.json 2> gpurun_out/r02_v2_sharded_2gpu_60k.err
echo "sharded_bench rc=$?"; tail -c 1500 gpurun_out/r02_v2_sharded_2gpu_60k.json
timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 \
    bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_v2_bench_2gpu.json 2> gpurun_out/r02_v2_bench_2gpu.err
echo "bench rc=$?"; tail -5 gpurun_out/r02_v2_bench_2gpu.err
