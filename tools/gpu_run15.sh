#!/bin/bash
# GPU run 15 (round 2, 1 GPU): chunked INT8-slice GEMMs for the triangular products of POTRI (opt-in) -- oracle parity
# at 8000 / 16 000 with the thresholds lowered, then the step probe at N = 50 000
mkdir -p gpurun_out
FVGP_OZAKI_LAUUM=2048 FVGP_OZAKI_ALL=1 FVGP_OZAKI_TRI=4 timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -s -k "16000 or 8000" > gpurun_out/r02_v15_pytest_ozaki_tri_parity.log 2>&1
echo "parity rc=$?"; tail -12 gpurun_out/r02_v15_pytest_ozaki_tri_parity.log
timeout 1200 python tools/ozaki_tri_probe.py > gpurun_out/r02_v15_ozaki_tri_probe.log 2>&1
echo "tri probe rc=$?"; tail -24 gpurun_out/r02_v15_ozaki_tri_probe.log
