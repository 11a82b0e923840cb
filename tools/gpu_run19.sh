#!/bin/bash
# GPU run 19 (round 2, 1 GPU): INT8-slice routing in the block-cyclic path (single rank), ncu --set full of one int8 GEMM
# launch, short bench with the int8 roofline record
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity_at_size.py -m gpu -q -k "sharded_single_rank or int8 or non_positive or api_dense_sharded" > gpurun_out/r02_v19_pytest_sharded_int8.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_v19_pytest_sharded_int8.log
timeout 600 ncu --set full --clock-control none -k regex:device_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_v19_i8_gemm \
  python tools/i8_one.py > gpurun_out/r02_v19_ncu_i8.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_v19_ncu_i8.log
ncu -i gpurun_out/r02_v19_i8_gemm.ncu-rep --page raw --csv > gpurun_out/r02_v19_i8_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_v19_i8_gemm.ncu-rep --page details > gpurun_out/r02_v19_i8_gemm_details.txt 2>/dev/null
ls -la gpurun_out/r02_v19_i8_gemm*
timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/r02_v19_bench.json 2> gpurun_out/r02_v19_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_v19_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds")}, d["e2e"]["value"], d["clocks"])
r = d["roofline"]
print({k: r[k] for k in ("kernel", "achieved", "peak", "frac", "frac_of_nominal", "int8_macs_per_step", "kernel_alone")})
print(r["fp64_equivalent"]["achieved"], r["fp64_equivalent"]["frac"], r["potrf"], r["potri"])
PY
