#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python tools/small_n_latency.py 2>&1 | grep -v Warn | tee gpurun_out/small_n_latency.log
python bench.py --workload c5 --size 20000 --grad --steps 1 --warmup 1 2> gpurun_out/bench_c5_small.err | tail -1 | tee gpurun_out/bench_c5_small.json; tail -3 gpurun_out/bench_c5_small.err
python bench.py --workload c5 --size 60000 --steps 1 --warmup 0 2> gpurun_out/bench_c5_1gpu.err | tail -1 | tee gpurun_out/bench_c5_1gpu.json; tail -3 gpurun_out/bench_c5_1gpu.err
