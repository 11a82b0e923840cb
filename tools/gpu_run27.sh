#!/bin/bash
# GPU run 27 (round 2, 1 GPU): ncu launch list (time only) of ONE timed C2 step of bench.py, final library
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 40000 --csv \
  --log-file gpurun_out/r02_v28_launches_bench_n50k.csv \
  python bench.py --steps 1 --warmup 1 --no-extras --no-parity --no-c4 --no-fresh-c4 > gpurun_out/r02_v28_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_v28_launches_bench_n50k.csv "python bench.py --steps 1 --warmup 1 --no-extras --no-parity --no-c4 --no-fresh-c4 (NVTX range timed/)" gpurun_out/r02_v28_launches_bench_n50k.json > gpurun_out/r02_v28_launches_bench_n50k.txt
head -24 gpurun_out/r02_v28_launches_bench_n50k.txt | cut -c1-160
rm -f gpurun_out/r02_v28_launches_bench_n50k.csv
