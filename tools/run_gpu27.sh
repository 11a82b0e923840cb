#!/bin/bash
# round-end validation on one B200: full GPU suite, smoke, bench lines (C2 headline, C1 population, C4), reference arm
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cat gpurun_out/bench_n50k.json
python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo c1 rc=$?; cat gpurun_out/bench_c1.json
python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; cut -c1-1200 gpurun_out/bench_c4_1m.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cut -c1-300 gpurun_out/bench_ref.json
FVGP_POTRF_TILE=1 python tools/small_n_latency.py 2>&1 | head -2 | sed 's/^/tile1 /'
python tools/small_n_latency.py 2>&1 | tail -6
