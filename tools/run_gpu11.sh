#!/bin/bash
# SpMV variant A/B on the C4 matrix, host profile of one C4 evaluation, bench K-fill timing fix
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spmv or gp2scale" 2>&1 | tail -5 | tee gpurun_out/pytest_sparse.log
for v in 0 1 2 3; do FVGP_SPMV_VARIANT=$v python tools/spmv_sweep.py 1000000; done 2>&1 | grep -v Warning | tee gpurun_out/spmv_sweep.log
FVGP_SPMV_VARIANT=1 python tools/spmv_sweep.py 150000 2>&1 | grep -E "spmv|rel err" | tee -a gpurun_out/spmv_sweep.log
python bench.py --workload c4 --steps 2 --warmup 1 --profile-host gpurun_out/c4_host_profile.txt > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json; head -60 gpurun_out/c4_host_profile.txt
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; tail -3 gpurun_out/bench_n50k.err; cat gpurun_out/bench_n50k.json
