#!/bin/bash
# GPU run 16 (round 2, 1 GPU): chunk count of the triangular INT8 products, block width of the INT8 POTRF
mkdir -p gpurun_out
PROBE_TRI=8,12,16,24,8,12,16,24 timeout 1200 python tools/ozaki_tri_probe.py > gpurun_out/r02_v16_ozaki_tri_probe.log 2>&1
echo "tri probe rc=$?"; tail -28 gpurun_out/r02_v16_ozaki_tri_probe.log
for nb in 2048 3072 4096; do
  FVGP_POTRF_NB=$nb timeout 300 python tools/potrf_nb_probe.py 2>&1 | grep -v "^\[fvgp" >> gpurun_out/r02_v16_potrf_nb_probe.log
done
cat gpurun_out/r02_v16_potrf_nb_probe.log
