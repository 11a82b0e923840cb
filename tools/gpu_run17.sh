#!/bin/bash
# GPU run 17 (round 2, 1 GPU): raw rate of the int8 GEMM per tile configuration
mkdir -p gpurun_out
timeout 600 python tools/i8_rate_probe.py > gpurun_out/r02_v17_i8_rate_probe.log 2>&1
echo "rc=$?"; cat gpurun_out/r02_v17_i8_rate_probe.log
