#!/bin/bash
# GPU run 21 (round 2, 1 GPU): ncu launch list (time only) of ONE timed C2 step of bench.py
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 40000 --csv \
  --log-file gpurun_out/r02_v21_launches_bench_n50k.csv \
  python bench.py --steps 1 --warmup 3 --no-extras --no-parity --no-c4 --no-fresh-c4 > gpurun_out/r02_v21_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_v21_bench_under_ncu.log | cut -c1-300
python tools/launch_summary.py gpurun_out/r02_v21_launches_bench_n50k.csv "python bench.py --steps 1 --warmup 3 --no-extras --no-parity --no-c4 --no-fresh-c4 (NVTX range timed/)" gpurun_out/r02_v21_launches_bench_n50k.json > gpurun_out/r02_v21_launches_bench_n50k.txt
head -40 gpurun_out/r02_v21_launches_bench_n50k.txt | cut -c1-200
rm -f gpurun_out/r02_v21_launches_bench_n50k.csv
