#!/bin/bash
# GPU run 26 (round 2, 1 GPU): scratch ladder / K-chunked LAUUM SYRK -- INT8 parity tests, N = 100 000 on one GPU, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "int8 or 16000" > gpurun_out/r02_v27_pytest_int8.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r02_v27_pytest_int8.log
timeout 400 python tools/c3_single_gpu_probe.py > gpurun_out/r02_v27_c3_single_gpu_probe.log 2>&1
echo "c3 probe rc=$?"; grep -v Warn gpurun_out/r02_v27_c3_single_gpu_probe.log | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --no-parity --no-fresh-c4 > gpurun_out/r02_v27_bench.json 2> gpurun_out/r02_v27_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v27_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"], d["roofline"]["potrf"]["seconds"], d["roofline"]["potri"]["seconds"])
a = d.get("int8_trailing_updates_ab", {}); print(a.get("lml_rel_diff_int8_vs_dmma"), a.get("grad_max_rel_diff_int8_vs_dmma"))
PY
