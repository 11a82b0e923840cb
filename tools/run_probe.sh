#!/bin/bash
# Runs every probe section in its own process so one faulting kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/probe_smi.txt 2>&1
for s in "environment" "fp64 peaks" "dgemm variants" "K-fill" "factorisation" "full dense" "gp2Scale" "timings"; do
  tag=$(echo $s | tr ' ' '_')
  PROBE_ONLY="$s" timeout 900 python tests/gpu_probe.py $@ > gpurun_out/probe_$tag.log 2>&1
  echo "section '$s' exit $?"
  tail -n 60 gpurun_out/probe_$tag.log | grep -E "FAIL|EXCEPTION|Error|error|failures" | head -20
done
