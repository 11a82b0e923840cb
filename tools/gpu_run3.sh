#!/bin/bash
# GPU run 3 (round 2, 2 GPUs): phase breakdown of the sharded gp2Scale step, NCCL sparse test, new tile kernel A/B
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "sparse or user_kernel" > gpurun_out/r02_v3_pytest_sharded.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_v3_pytest_sharded.log
for T in 1 0; do
FVGP_SHARDED_TIMING=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 \
    bench.py --workload c4 --sharded --steps 5 --warmup 2 > gpurun_out/r02_v3_c4_sharded_2gpu_timing$T.json 2> gpurun_out/r02_v3_c4_sharded_2gpu_timing$T.err
echo "c4 sharded timing=$T rc=$?"
done
# single-GPU: the new tile kernel against the previous one (POTRF at N = 1000 / 4096 / 8192 / 16384, LML latency), full parity suite
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regressions.py -m gpu -x -q > gpurun_out/r02_v3_pytest_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r02_v3_pytest_parity.log
for TILE in 3 2; do
FVGP_POTRF_TILE=$TILE CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/potrf_probe.py > gpurun_out/r02_v3_potrf_tile$TILE.log 2>&1
echo "potrf tile=$TILE rc=$?"; tail -12 gpurun_out/r02_v3_potrf_tile$TILE.log
done
