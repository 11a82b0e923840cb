#!/bin/bash
# GPU run 10 (round 2, 1 GPU): full -m gpu suite, Ozaki one-stream step probe, POTRF vs cuSOLVER incl. N = 50 000
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_v10_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/r02_v10_pytest_all.log
timeout 900 python tools/ozaki_step_probe.py > gpurun_out/r02_v10_ozaki_step_probe.log 2>&1
echo "ozaki step probe rc=$?"; tail -17 gpurun_out/r02_v10_ozaki_step_probe.log
timeout 600 python tools/potrf_probe.py --big 2>&1 | grep "N=50000\|N=16384" > gpurun_out/r02_v10_potrf_vs_cusolver_50k.log; cat gpurun_out/r02_v10_potrf_vs_cusolver_50k.log
