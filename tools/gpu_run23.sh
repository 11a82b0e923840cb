#!/bin/bash
# GPU run 23 (round 2, 1 GPU): vectorised slice kernel -- accuracy probe, INT8 parity, POTRF probe, short bench
mkdir -p gpurun_out
timeout 600 python tools/ozaki_probe.py > gpurun_out/r02_v23_ozaki_probe.log 2>&1
echo "probe rc=$?"; tail -8 gpurun_out/r02_v23_ozaki_probe.log
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_sharded.py -m gpu -q -k "int8" > gpurun_out/r02_v23_pytest_int8.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r02_v23_pytest_int8.log
timeout 300 python tools/potrf_nb_probe.py 2>&1 | grep -v "^\[fvgp"
timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/r02_v23_bench.json 2> gpurun_out/r02_v23_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v23_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"], d["clocks"])
r = d["roofline"]
print({k: r[k] for k in ("achieved", "peak", "frac", "frac_of_nominal", "int8_macs_per_step")}, r["potrf"]["seconds"], r["potri"]["seconds"])
print(d.get("int8_trailing_updates_ab"))
PY
