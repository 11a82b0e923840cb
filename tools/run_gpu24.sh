#!/bin/bash
# lock-step population schedule (batched kernels): GPU suite, C1 population bench A/B against the stream schedule
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
for s in 0 1; do
  FVGP_POPULATION_STREAMS=$s timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.streams$s.json 2> gpurun_out/bench_c1.streams$s.err; echo "c1 streams=$s rc=$?"; cut -c1-1400 gpurun_out/bench_c1.streams$s.json; tail -3 gpurun_out/bench_c1.streams$s.err
done
timeout 300 python bench.py --workload c1 --population 200 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1_pop200.json 2> gpurun_out/bench_c1_pop200.err; echo c1-200 rc=$?; cut -c1-1400 gpurun_out/bench_c1_pop200.json; tail -3 gpurun_out/bench_c1_pop200.err
for s in 0 1; do
FVGP_POPULATION_STREAMS=$s timeout 300 python bench.py --workload c1 --size 4000 --population 16 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1_n4000.streams$s.json 2> gpurun_out/bench_c1_n4000.err; echo c1-4000 streams=$s rc=$?; cut -c1-1400 gpurun_out/bench_c1_n4000.streams$s.json; tail -3 gpurun_out/bench_c1_n4000.err
done
