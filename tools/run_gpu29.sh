#!/bin/bash
# batched population fill: full GPU suite, smoke, C1 population line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?; cat gpurun_out/bench_c1.json; tail -2 gpurun_out/bench_c1.err
python bench.py --workload c1 --population 200 --no-cpu-baseline > gpurun_out/bench_c1_pop200.json 2> gpurun_out/bench_c1_pop200.err; echo c1-200 rc=$?; cut -c1-200 gpurun_out/bench_c1_pop200.json
