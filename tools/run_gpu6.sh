#!/bin/bash
mkdir -p gpurun_out
for s in "K-fill" "timings"; do
  tag=$(echo $s | tr ' ' '_')
  PROBE_ONLY="$s" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_$tag.log 2>&1
  echo "section '$s' exit $?"; grep -E "FAIL|EXCEPTION|Error" gpurun_out/probe_$tag.log | head -10
done
grep -E "kfill|fill|LML" gpurun_out/probe_timings.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
