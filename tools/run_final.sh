#!/bin/bash
# round-end validation on one B200: tests, smoke, probes, bench (C2 + C4 + reference arm), ncu launch list of ONE timed C2 step
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
grep -q "failed" gpurun_out/pytest_gpu.log && { echo "TESTS FAILED - stopping"; cat gpurun_out/pytest_gpu.log; exit 1; }
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for s in "K-fill" "timings" "fp64 peaks"; do
  tag=$(echo $s | tr ' ' '_')
  PROBE_ONLY="$s" timeout 900 python tests/gpu_probe.py > gpurun_out/probe_$tag.log 2>&1
  echo "section '$s' exit $?"; grep -E "FAIL|EXCEPTION|Error" gpurun_out/probe_$tag.log | head -10
done
grep -E "kfill|fill|LML|wendland|spmv|pcg" gpurun_out/probe_timings.log; grep -E "potrf|potri|gemm|peak" gpurun_out/probe_fp64_peaks.log | head -20
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n50k.json 2> gpurun_out/bench_n50k.err; echo bench rc=$?; cat gpurun_out/bench_n50k.json
python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_c4_1m.json 2> gpurun_out/bench_c4_1m.err; echo c4 rc=$?; cat gpurun_out/bench_c4_1m.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --impl reference --workload c4 --steps 1 --warmup 0 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err; echo ref c4 rc=$?; cut -c1-600 gpurun_out/bench_ref_c4.json
if [ "$1" = "ncu" ]; then
  # only the kernels inside bench.py's NVTX range "timed" (one LML + gradient evaluation) are profiled
  timeout 1500 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo ncu rc=$?
  python tools/launch_summary.py /tmp/launches_c2.csv "python bench.py --steps 1 --warmup 1 --no-cpu-baseline (N=50000), NVTX range 'timed': every kernel of ONE timed evaluation (LML + gradient)" > gpurun_out/launches_bench_n50k.txt; head -24 gpurun_out/launches_bench_n50k.txt
fi
