#!/bin/bash
mkdir -p gpurun_out
for s in "fp64 peaks" "K-fill" "factorisation" "full dense" "timings"; do
  tag=$(echo $s | tr ' ' '_')
  PROBE_ONLY="$s" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_$tag.log 2>&1
  echo "section '$s' exit $?"; grep -E "FAIL|EXCEPTION|Error" gpurun_out/probe_$tag.log | head -10
done
grep -E "dgemm|syrk|potrf|potri" gpurun_out/probe_fp64_peaks.log; grep -E "kfill|LML|wendland|spmv|pcg" gpurun_out/probe_timings.log
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_n50k.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_seconds_per_step'], d['roofline_kfill']['achieved'], d['gpu_launches'])"
