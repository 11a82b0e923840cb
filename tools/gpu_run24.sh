#!/bin/bash
# GPU run 24 (round 2, N GPUs = $1): the N-rank bench line as the driver launches it (replicas headline + sharded sub-records)
N=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02_v24_bench_${N}gpu.json 2> gpurun_out/r02_v24_bench_${N}gpu.err
echo "bench N=$N rc=$?"; grep "^\[bench" gpurun_out/r02_v24_bench_${N}gpu.err | sort -u | tail -12
python - <<'PY'
import json, sys, glob
for f in sorted(glob.glob("gpurun_out/r02_v24_bench_*gpu.json")):
    lines = [l for l in open(f) if l.startswith("{")]
    if not lines:
        print(f, "no JSON line"); continue
    d = json.loads(lines[-1])
    sh = d.get("sharded", {})
    print(f, "value", d["value"], "wall", d.get("wall_seconds"), d.get("clocks"))
    for k, v in sh.items():
        if isinstance(v, dict):
            print(" ", k, {a: v.get(a) for a in ("n", "n_gpus", "grid", "seconds_per_step", "tflops_per_gpu", "strong_scaling_efficiency", "value", "ms_per_step", "error", "phase_seconds_last_step")})
            for kk in ("single_gpu", "single_gpu_dmma_only"):
                if v.get(kk):
                    print("    ", kk, {a: v[kk].get(a) for a in ("seconds_per_step", "lml_rel_diff", "grad_rel_diff", "agree_1e-8")})
PY
