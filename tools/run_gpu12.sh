#!/bin/bash
# lanes=columns SpMM A/B, rounded CSR allocations, full GPU test pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for v in 0 1; do FVGP_LANCZOS_COLS=$v python tools/spmv_sweep.py 1000000 2>&1 | grep -E "slq|spmv" | sed "s/^/cols=$v /"; done | tee gpurun_out/slq_sweep.log
python bench.py --workload c4 --steps 3 --warmup 2 --profile-host gpurun_out/c4_host_profile.txt > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json; head -24 gpurun_out/c4_host_profile.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"spmv_kernel" -c 1 -o gpurun_out/ncu_spmv python tests/ncu_targets.py spmv 1000000 > gpurun_out/ncu_spmv.log 2>&1; echo ncu spmv rc=$?
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"lanczos_spmm" -c 1 -o gpurun_out/ncu_slq python tests/ncu_targets.py slq 1000000 > gpurun_out/ncu_slq.log 2>&1; echo ncu slq rc=$?
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"kfill_kernel" -c 1 -o gpurun_out/ncu_kfill python tests/ncu_targets.py kfill 50000 > gpurun_out/ncu_kfill.log 2>&1; echo ncu kfill rc=$?
