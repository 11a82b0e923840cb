#!/bin/bash
# GPU run 7 (round 2, 1 GPU): host profile of the C4 step (regression hunt), Ozaki step probe, POTRS check, sharded non-PD test
mkdir -p gpurun_out
timeout 600 python tools/c4_host_profile.py > gpurun_out/r02_v7_c4_host_profile.log 2>&1
echo "c4 profile rc=$?"; head -50 gpurun_out/r02_v7_c4_host_profile.log
FVGP_SLQ_GREEDY=1 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r02_v7_c4_greedy.json 2> gpurun_out/r02_v7_c4_greedy.err
echo "c4 greedy rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_v7_c4_greedy.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['phases_seconds'])"
timeout 900 python tools/ozaki_step_probe.py > gpurun_out/r02_v7_ozaki_step_probe.log 2>&1
echo "ozaki step probe rc=$?"; tail -20 gpurun_out/r02_v7_ozaki_step_probe.log
timeout 300 python tools/potrf_probe.py 2>&1 | grep "N=8192\|N=16384" > gpurun_out/r02_v7_potrs_check.log; cat gpurun_out/r02_v7_potrs_check.log
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r02_v7_pytest_sharded_1gpu.log 2>&1; tail -3 gpurun_out/r02_v7_pytest_sharded_1gpu.log
