#!/bin/bash
# population evaluator + second-generation tile kernel: GPU suite, small-N latency A/B, C1 population bench, potrf timing at 8192 / 50k
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
for v in 1 0; do
  echo "== FVGP_POTRF_TILE=$v"; FVGP_POTRF_TILE=$v timeout 300 python tools/small_n_latency.py 2>&1 | tail -6 | tee gpurun_out/small_n_latency.tile$v.log
done
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?; cat gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
timeout 300 python bench.py --workload c1 --size 4000 --population 16 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1_n4000.json 2> gpurun_out/bench_c1_n4000.err; echo c1-4000 rc=$?; cat gpurun_out/bench_c1_n4000.json
for v in 1 0; do
  FVGP_POTRF_TILE=$v timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n50k.tile$v.json 2> gpurun_out/bench_n50k.tile$v.err; echo "tile$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_n50k.tile$v.json"))
print("tile$v", d["value"], d["roofline"]["frac"], d["roofline"]["phase_seconds_per_step"])
PY
done
