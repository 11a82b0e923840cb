#!/bin/bash
# GPU run 6 (round 2, N GPUs = $1): the N-rank bench line (replicas headline + sharded sub-records) as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_v6_bench_${N}gpu.json 2> gpurun_out/r02_v6_bench_${N}gpu.err
echo "bench N=$N rc=$?"; grep "^\[bench" gpurun_out/r02_v6_bench_${N}gpu.err | sort -u | tail -12
python - <<'PY'
import json, sys, glob
for f in sorted(glob.glob("gpurun_out/r02_v6_bench_*gpu.json")):
    lines = [l for l in open(f) if l.startswith("{")]
    if not lines:
        print(f, "no JSON line"); continue
    d = json.loads(lines[-1])
    sh = d.get("sharded", {})
    print(f, "value", d["value"], "wall", d.get("wall_seconds"))
    for k, v in sh.items():
        if isinstance(v, dict):
            print(" ", k, {a: v.get(a) for a in ("n", "n_gpus", "grid", "seconds_per_step", "tflops_per_gpu", "strong_scaling_efficiency", "value", "ms_per_step", "error")},
                  (v.get("single_gpu") or {}).get("agree_1e-8"))
PY
