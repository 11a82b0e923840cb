#!/bin/bash
# GPU run 9 (round 2, 2 GPUs): why is the sharded gp2Scale step slower inside the full bench line than standalone?
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
FVGP_SHARDED_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 2 --steps 1 --warmup 1 --c3-points 4000 > gpurun_out/r02_v9_bench_2gpu_timing.json 2> gpurun_out/r02_v9_bench_2gpu_timing.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 \
    bench.py --workload c4 --sharded --steps 5 --warmup 2 > gpurun_out/r02_v9_c4_sharded_standalone.json 2> gpurun_out/r02_v9_c4_sharded_standalone.err
echo "c4 standalone rc=$?"
FVGP_SHARDED_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 \
    bench.py --workload c4 --sharded --steps 5 --warmup 2 > gpurun_out/r02_v9_c4_sharded_standalone_timing.json 2> gpurun_out/r02_v9_c4_sharded_standalone_timing.err
python - <<'PY'
import json
for f in ("r02_v9_bench_2gpu_timing", "r02_v9_c4_sharded_standalone", "r02_v9_c4_sharded_standalone_timing"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        c = d.get("sharded", {}).get("c4_gp2scale_sharded", d)
        print(f, c.get("ms_per_step"), c.get("phase_ms_per_evaluation"), c.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
# single GPU: C4 with / without the SLQ overlap
for OV in 1 0; do
FVGP_SLQ_OVERLAP=$OV CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/r02_v9_c4_overlap$OV.json 2> gpurun_out/r02_v9_c4_overlap$OV.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_v9_c4_overlap$OV.json') if l.startswith('{')][-1]); print('overlap=$OV', d['value'], d['ms_per_step'], d['last_lml'], d['phases_seconds'])"
done
