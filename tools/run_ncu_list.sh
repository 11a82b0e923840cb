#!/bin/bash
mkdir -p gpurun_out
timeout 1700 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/bench_under_ncu.log | cut -c1-300
python tools/launch_summary.py /tmp/launches_c2.csv "python bench.py --steps 1 --warmup 1 --no-cpu-baseline (N=50000), NVTX range 'timed': every kernel of ONE timed evaluation (LML + gradient)" > gpurun_out/launches_bench_n50k.txt; head -24 gpurun_out/launches_bench_n50k.txt
