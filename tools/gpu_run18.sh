#!/bin/bash
# GPU run 18 (round 2, 1 GPU): 256 x 256 x 128 CTA-pair tile as the default of the int8 GEMM -- POTRF / POTRI probes,
# the new INT8 parity test, short bench
mkdir -p gpurun_out
for nb in 2048 4096; do
  FVGP_POTRF_NB=$nb timeout 300 python tools/potrf_nb_probe.py 2>&1 | grep -v "^\[fvgp" >> gpurun_out/r02_v18_potrf_nb_probe.log
done
cat gpurun_out/r02_v18_potrf_nb_probe.log
PROBE_TRI=0,4,8,12,16 timeout 1200 python tools/ozaki_tri_probe.py > gpurun_out/r02_v18_ozaki_tri_probe.log 2>&1
echo "tri probe rc=$?"; grep -v "^\[fvgp" gpurun_out/r02_v18_ozaki_tri_probe.log | tail -24
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "int8 or 16000 or benchmarked_n" > gpurun_out/r02_v18_pytest_int8_parity.log 2>&1
echo "parity rc=$?"; tail -6 gpurun_out/r02_v18_pytest_int8_parity.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/r02_v18_bench.json 2> gpurun_out/r02_v18_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v18_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"], d["roofline"]["potrf"]["seconds"], d["roofline"]["potri"])
print(d.get("int8_trailing_updates_ab"))
PY
