#!/bin/bash
# GPU run 4 (round 2, 1 GPU): N=1 bench line (crash hunt with faulthandler), K-fill mirror A/B, parity suite on the new
# fill / Lanczos paths, ncu launch list of POTRF at N = 8192
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_plugin.py tests/test_gpu_parity_at_size.py -k 'not fifty and not benchmarked_n and not 16000' -m gpu -q > gpurun_out/r02_v4_pytest_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r02_v4_pytest_parity.log
for B in 1 2; do timeout 300 python tools/kfill_sweep.py --bulk $B 30000 50000 >> gpurun_out/r02_v4_kfill_sweep.log 2>&1; done
cat gpurun_out/r02_v4_kfill_sweep.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_v4_bench.json 2> gpurun_out/r02_v4_bench.err
echo "bench rc=$?"; grep -v Warning gpurun_out/r02_v4_bench.err | tail -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_v4_launches_potrf_8192.csv \
    python tools/potrf_sweep.py 8192 > gpurun_out/r02_v4_ncu_potrf.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_v4_ncu_potrf.log
# INT8-slice (Ozaki) GEMM: accuracy + throughput probe, POTRF at N = 50 000 with and without it, LML / gradient parity with it
timeout 600 python tools/ozaki_probe.py > gpurun_out/r02_v4_ozaki_probe.log 2>&1
echo "ozaki probe rc=$?"; tail -30 gpurun_out/r02_v4_ozaki_probe.log
for OZ in 0 8; do
FVGP_OZAKI=$OZ timeout 300 python tools/potrf_sweep.py 50000 > gpurun_out/r02_v4_potrf_50k_ozaki$OZ.log 2>&1
echo "potrf 50k ozaki=$OZ rc=$?"; tail -2 gpurun_out/r02_v4_potrf_50k_ozaki$OZ.log
done
FVGP_OZAKI=8 timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "benchmarked_n or 16000" > gpurun_out/r02_v4_pytest_ozaki_parity.log 2>&1
echo "ozaki parity rc=$?"; tail -5 gpurun_out/r02_v4_pytest_ozaki_parity.log
