#!/bin/bash
# tile kernel with the shortened per-column chain: GPU suite, smoke, latency tool, C1 + C2 bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/small_n_latency.py 2>&1 | tail -5
python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?; cut -c1-120 gpurun_out/bench_c1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c1.json')); print({k:d[k] for k in ('value','ms_per_step_median','device_ms_per_step','gpu_launches','one_at_a_time','population_with_gradient')})"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_n50k.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_seconds_per_step'], d['clocks'])"
