#!/bin/bash
# re-entry validation on one B200: GPU suite, smoke, C2 bench line (short)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cat gpurun_out/bench_n50k.json
