#!/bin/bash
# round-1 re-entry: tests, bench (C2 + C4), timings probe, ncu launch list of the bench command
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
PROBE_ONLY="timings" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_timings.log 2>&1; grep -E "kfill|fill|LML|wendland|spmv|pcg" gpurun_out/probe_timings.log
PROBE_ONLY="fp64 peaks" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_fp64_peaks.log 2>&1; grep -E "potrf|potri|gemm|peak" gpurun_out/probe_fp64_peaks.log | head -30
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cat gpurun_out/bench_n50k.json
python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cat gpurun_out/bench_ref.json
