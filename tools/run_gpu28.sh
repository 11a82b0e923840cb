#!/bin/bash
# fused small-N substitution kernel: full GPU suite, A/B on the C1 population bench and the small-N latency tool, C2 line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for f in 0 1; do
  FVGP_TRSV_FUSED=$f python bench.py --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.fused$f.json 2> gpurun_out/bench_c1.fused$f.err; echo "c1 fused=$f rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c1.fused$f.json"))
print("fused=$f", "evals/s", round(d["value"]), "median ms", round(d["ms_per_step_median"],2), "device ms", round(d["device_ms_per_step"],2), "launches", d["gpu_launches"], "grad", round(d["population_with_gradient"]["value"]), "single", round(d["one_at_a_time"]["ms_per_eval"],3))
PY
done
FVGP_TRSV_FUSED=0 python tools/small_n_latency.py 2>&1 | head -2 | sed 's/^/steps /'
python tools/small_n_latency.py 2>&1 | tail -6
python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cut -c1-900 gpurun_out/bench_n50k.json
