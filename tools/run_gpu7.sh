#!/bin/bash
# re-entry validation of HEAD: tests, smoke, probes, bench (C2 + C4 + reference), ncu launch list of the bench command
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for s in "K-fill" "timings" "fp64 peaks"; do
  tag=$(echo $s | tr ' ' '_')
  PROBE_ONLY="$s" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_$tag.log 2>&1
  echo "section '$s' exit $?"; grep -E "FAIL|EXCEPTION|Error" gpurun_out/probe_$tag.log | head -10
done
grep -E "kfill|fill|LML|wendland|spmv|pcg" gpurun_out/probe_timings.log; grep -E "potrf|potri|gemm|peak" gpurun_out/probe_fp64_peaks.log | head -30
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cat gpurun_out/bench_n50k.json
python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; tail -3 gpurun_out/bench_c4_1m.err; cat gpurun_out/bench_c4_1m.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo ncu rc=$?
python tools/launch_summary.py /tmp/launches_c2.csv "python bench.py --steps 1 --warmup 1 --no-cpu-baseline (N=50000; constructor + 1 warm-up + 1 timed + 1 e2e evaluation)" > gpurun_out/launches_bench_n50k.txt; head -20 gpurun_out/launches_bench_n50k.txt
