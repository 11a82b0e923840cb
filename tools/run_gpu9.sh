#!/bin/bash
# verify: K-fill rounding fix, chunked Wendland units, SpMV gather fix, blocked potrs, potrf width heuristic; K-fill placement experiments
mkdir -p gpurun_out
python tools/debug_kfill.py > gpurun_out/debug_kfill.log 2>&1; cat gpurun_out/debug_kfill.log
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
python tools/kfill_sweep.py 50000 2>&1 | tee gpurun_out/kfill_sweep.log
python tools/kfill_sweep.py 50000 --hog 48 2>&1 | tee -a gpurun_out/kfill_sweep.log
python tools/kfill_sweep.py 50000 --after-gemm 2>&1 | tee -a gpurun_out/kfill_sweep.log
python tools/potrf_sweep.py 50000 --potri | tail -1 | tee gpurun_out/potrf_sweep.log
python tools/potrf_sweep.py 16384 --potri | tail -1 | tee -a gpurun_out/potrf_sweep.log
python tools/potrf_sweep.py 8192 --potri | tail -1 | tee -a gpurun_out/potrf_sweep.log
FVGP_POTRF_NB=3072 python tools/potrf_sweep.py 50000 | tail -1 | tee -a gpurun_out/potrf_sweep.log
FVGP_POTRF_NB=1536 python tools/potrf_sweep.py 50000 | tail -1 | tee -a gpurun_out/potrf_sweep.log
PROBE_ONLY="timings" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_timings.log 2>&1; grep -E "LML|wendland|spmv|pcg|FAIL|Error" gpurun_out/probe_timings.log
python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; tail -3 gpurun_out/bench_n50k.err; cat gpurun_out/bench_n50k.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"wendland_csr_kernel" -c 2 -o gpurun_out/ncu_wendland python tests/ncu_targets.py wendland 400000 > gpurun_out/ncu_wendland.log 2>&1; echo ncu wendland rc=$?
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"spmv_kernel" -c 1 -o gpurun_out/ncu_spmv python tests/ncu_targets.py spmv 400000 > gpurun_out/ncu_spmv.log 2>&1; echo ncu spmv rc=$?
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"lanczos_spmm" -c 1 -o gpurun_out/ncu_slq python tests/ncu_targets.py slq 400000 > gpurun_out/ncu_slq.log 2>&1; echo ncu slq rc=$?
