#!/bin/bash
# GPU run 8 (round 2, 2 GPUs): NCCL tests after the round-robin diagonal inversions / no-sync panel POTRF, 2-rank bench line
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_v8_pytest_sharded.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_v8_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    tests/sharded_bench.py --size 60000 --grad --check > gpurun_out/r02_v8_sharded_2gpu_60k.json 2> gpurun_out/r02_v8_sharded_2gpu_60k.err
echo "sharded_bench rc=$?"; tail -c 1200 gpurun_out/r02_v8_sharded_2gpu_60k.json
bash tools/gpu_run6.sh 2
mv gpurun_out/r02_v6_bench_2gpu.json gpurun_out/r02_v8_bench_2gpu.json; mv gpurun_out/r02_v6_bench_2gpu.err gpurun_out/r02_v8_bench_2gpu.err
