#!/bin/bash
# GPU run 12 (round 2, 1 GPU): INT8 GEMM tile 1-SM vs 2-SM, one vs two streams, new look-ahead widths, headline with the defaults
mkdir -p gpurun_out
for T in 1 2; do
FVGP_OZAKI_TILE=$T timeout 600 python tools/ozaki_probe.py 2>&1 | grep "SYRK\|ozaki slices=8\|slices=8:" > gpurun_out/r02_v12_ozaki_probe_tile$T.log
echo "tile=$T"; cat gpurun_out/r02_v12_ozaki_probe_tile$T.log
done
for T in 1 2; do for ST in 1 2; do
FVGP_OZAKI_TILE=$T FVGP_OZAKI_STREAMS=$ST timeout 300 python tools/potrf_sweep.py 50000 2>&1 | tail -1 | sed "s/^/tile=$T streams=$ST /" >> gpurun_out/r02_v12_potrf_50k_ozaki_variants.log
done; done
cat gpurun_out/r02_v12_potrf_50k_ozaki_variants.log
timeout 600 python tools/potrf_probe.py > gpurun_out/r02_v12_potrf_probe.log 2>&1; grep "^N=" gpurun_out/r02_v12_potrf_probe.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/r02_v12_bench.json 2> gpurun_out/r02_v12_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v12_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "dmma_kernel_only", "whole_step")}, d["roofline"]["potrf"]["seconds"])
print(d.get("int8_trailing_updates_ab"))
print("c4", d["c4"]["value"], d["c4"]["ms_per_step"], "c1", d["c1"]["value"])
PY
