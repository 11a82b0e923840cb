#!/bin/bash
# final check: GPU suite incl. the n = 6200 population case (stream schedule + look-ahead POTRF per stream); population at N = 8000
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload c1 --size 8000 --population 8 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1_n8000.json 2> gpurun_out/bench_c1_n8000.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_c1_n8000.json')); print({k:d[k] for k in ('value','ms_per_step_median','device_ms_per_step','gpu_launches','one_at_a_time','population_with_gradient')})"
