#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
python tools/append_timing.py 2>&1 | grep -v Warn | tee gpurun_out/append_timing.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
