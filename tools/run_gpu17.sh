#!/bin/bash
# validate the 64x64-tile GEMM variant and the micro-tiled trace kernel; A/B small tiles on/off
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
PROBE_ONLY="fp64 peaks" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_fp64_peaks.log 2>&1; grep -E "potrf|potri|gemm|FAIL|Error" gpurun_out/probe_fp64_peaks.log | head -20
PROBE_ONLY="factorisation" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_factorisation.log 2>&1; grep -E "FAIL|Error|EXCEPTION" gpurun_out/probe_factorisation.log | head; tail -3 gpurun_out/probe_factorisation.log
for v in 0 1; do for n in 2048 8192 16384; do FVGP_GEMM_SMALL_TILES=$v python tools/potrf_sweep.py $n --potri | tail -1 | sed "s/^/small_tiles=$v /"; done; done | tee gpurun_out/small_tiles_sweep.log
python tools/potrf_sweep.py 50000 --potri | tail -1 | tee -a gpurun_out/small_tiles_sweep.log
python tools/small_n_latency.py 2>&1 | grep -v Warn | tee gpurun_out/small_n_latency.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; tail -3 gpurun_out/bench_n50k.err; cat gpurun_out/bench_n50k.json
