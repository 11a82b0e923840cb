#!/bin/bash
# C1 population bench: cost of the clock sampler itself (in-process NVML vs forked nvidia-smi vs none); new kernel names
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "robust or population or fused" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_partial.log
for m in nvml none smi nvml; do
  FVGP_BENCH_SAMPLER=$m timeout 200 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.$m.json 2> gpurun_out/bench_c1.$m.err; echo "sampler=$m rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c1.$m.json"))
print("$m", "evals/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "device ms", round(d["device_ms_per_step"],2), "grad", round(d["population_with_gradient"]["value"]), d["clocks"])
PY
done
