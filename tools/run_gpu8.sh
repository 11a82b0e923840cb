#!/bin/bash
# validation of the sparse rewrite, look-ahead potrf, banded fill order; K-fill mode mismatch debug
mkdir -p gpurun_out
python tools/debug_kfill.py > gpurun_out/debug_kfill.log 2>&1; cat gpurun_out/debug_kfill.log
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
for b in 0 4 8 16; do FVGP_FILL_BAND=$b python tools/kfill_sweep.py 30000 50000; done 2>&1 | tee gpurun_out/kfill_sweep.log
for nb in 0 1024 2048 4096; do FVGP_POTRF_NB=$nb python tools/potrf_sweep.py 50000 | tail -1; done 2>&1 | tee gpurun_out/potrf_sweep.log
FVGP_POTRF_NB=2048 python tools/potrf_sweep.py 16384 | tail -1 | tee -a gpurun_out/potrf_sweep.log
FVGP_POTRF_NB=1024 python tools/potrf_sweep.py 16384 | tail -1 | tee -a gpurun_out/potrf_sweep.log
PROBE_ONLY="timings" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_timings.log 2>&1; grep -E "wendland|spmv|pcg|FAIL|Error" gpurun_out/probe_timings.log
python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; tail -3 gpurun_out/bench_n50k.err; cat gpurun_out/bench_n50k.json
for t in wendland spmv slq; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:"wendland_csr|spmv_kernel|lanczos_spmm" -c 2 -o gpurun_out/ncu_$t python tests/ncu_targets.py $t 400000 > gpurun_out/ncu_$t.log 2>&1; echo ncu $t rc=$?
done
