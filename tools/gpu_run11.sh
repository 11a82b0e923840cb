#!/bin/bash
# GPU run 11 (round 2, 1 GPU): Ozaki POTRF with the scratch kept in the stream-ordered pool; look-ahead block width sweep at mid N
mkdir -p gpurun_out
PROBE_N=50000 timeout 900 python tools/ozaki_step_probe.py > gpurun_out/r02_v11_ozaki_step_probe.log 2>&1
echo "ozaki step probe rc=$?"; tail -17 gpurun_out/r02_v11_ozaki_step_probe.log
for NB in 0 512 1024 2048; do for N in 8192 12288 16384 24576; do
FVGP_POTRF_NB=$NB timeout 120 python tools/potrf_sweep.py $N 2>&1 | tail -1 >> gpurun_out/r02_v11_potrf_nb_sweep.log
done; done
cat gpurun_out/r02_v11_potrf_nb_sweep.log
