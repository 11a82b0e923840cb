#!/bin/bash
# GPU run 22 (round 2, 1 GPU): INT8 look-ahead column update in POTRF (A/B), smaller recursion levels of POTRI on the INT8 path
mkdir -p gpurun_out
for p in 0 1; do
  FVGP_OZAKI_PANEL=$p timeout 300 python tools/potrf_nb_probe.py 2>&1 | grep -v "^\[fvgp" | sed "s/^/panel_int8=$p /" >> gpurun_out/r02_v22_potrf_panel_probe.log
done
cat gpurun_out/r02_v22_potrf_panel_probe.log
for mr in 8192 4096 2048; do
  PROBE_MIN_ROWS=$mr PROBE_TRI=8,8 timeout 600 python tools/ozaki_tri_probe.py 2>&1 | grep -v "^\[fvgp" | grep -v "tri=0" >> gpurun_out/r02_v22_potri_min_rows_probe.log
done
cat gpurun_out/r02_v22_potri_min_rows_probe.log
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "int8 or benchmarked_n" > gpurun_out/r02_v22_pytest_int8_parity.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_v22_pytest_int8_parity.log
