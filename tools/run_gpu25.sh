#!/bin/bash
# ncu evidence of this session's work: launch list (time + DRAM bytes) of ONE C2 evaluation and of ONE C1 population
# step, and a full capture of the blocked tile kernel; the C1 bench line with the device-time split
mkdir -p gpurun_out
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?; cut -c1-1500 gpurun_out/bench_c1.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics $M --clock-control none --csv --log-file /tmp/launches_c1.csv python bench.py --workload c1 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1_under_ncu.log 2>&1; echo ncu c1 rc=$?
python tools/launch_summary.py /tmp/launches_c1.csv "python bench.py --workload c1 --steps 1 --warmup 2 --no-cpu-baseline (N=1000, 40 proposals), NVTX range 'timed': every kernel of ONE population step" gpurun_out/launches_c1_population.json > gpurun_out/launches_c1_population.txt; head -30 gpurun_out/launches_c1_population.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"potrf_tile2" -c 1 -o gpurun_out/ncu_potrf_tile2 python bench.py --workload c1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_potrf_tile2.log 2>&1; echo ncu tile rc=$?
python profiles/summarize_ncu.py gpurun_out/ncu_potrf_tile2.ncu-rep > gpurun_out/ncu_potrf_tile2.summary.txt; cat gpurun_out/ncu_potrf_tile2.summary.txt
timeout 1500 ncu --nvtx --nvtx-include "timed/" --metrics $M --clock-control none --csv --log-file /tmp/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo ncu c2 rc=$?
python tools/launch_summary.py /tmp/launches_c2.csv "python bench.py --steps 1 --warmup 1 --no-cpu-baseline (N=50000), NVTX range 'timed': every kernel of ONE timed evaluation (LML + gradient)" gpurun_out/launches_bench_n50k.json > gpurun_out/launches_bench_n50k.txt; head -24 gpurun_out/launches_bench_n50k.txt
