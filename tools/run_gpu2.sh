#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
PROBE_ONLY="fp64 peaks" python tests/gpu_probe.py --quick > gpurun_out/probe_fp64_peaks.log 2>&1; grep -E "dgemm|syrk|potrf|potri" gpurun_out/probe_fp64_peaks.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:dgemm_mma -s 1 -c 1 -f -o gpurun_out/ncu_gemm python tests/ncu_targets.py gemm 8192 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
$NCU -k regex:kfill_kernel -s 1 -c 1 -f -o gpurun_out/ncu_kfill python tests/ncu_targets.py kfill 30000 > gpurun_out/ncu_kfill.log 2>&1; echo "ncu kfill rc=$?"
$NCU -k regex:kfill_kernel -s 1 -c 1 -f -o gpurun_out/ncu_kfill_full python tests/ncu_targets.py kfill_full 30000 > gpurun_out/ncu_kfill_full.log 2>&1; echo "ncu kfill_full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_n8192.csv python bench.py --n 8192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | head -30
