#!/bin/bash
# GPU run 13 (round 2, 1 GPU): INT8 SYRK inside LAUUM (opt-in) -- step probe at N = 50 000, oracle parity at 16 000 / 50 000
mkdir -p gpurun_out
FVGP_OZAKI_LAUUM=1 timeout 900 python tools/ozaki_step_probe.py > gpurun_out/r02_v13_ozaki_lauum_step_probe.log 2>&1
echo "lauum step probe rc=$?"; tail -17 gpurun_out/r02_v13_ozaki_lauum_step_probe.log
FVGP_OZAKI_LAUUM=4096 FVGP_OZAKI_ALL=1 timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "benchmarked_n or 16000 or 8000" > gpurun_out/r02_v13_pytest_ozaki_lauum_parity.log 2>&1
echo "parity rc=$?"; tail -6 gpurun_out/r02_v13_pytest_ozaki_lauum_parity.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/r02_v13_bench.json 2> gpurun_out/r02_v13_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v13_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"], d["roofline"]["potrf"]["seconds"])
print("c4", d["c4"]["value"], d["c4"]["ms_per_step"], "fresh", d.get("c4_fresh_process"))
PY
