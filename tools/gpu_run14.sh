#!/bin/bash
# GPU run 14 (round 2, 1 GPU): what the driver runs at round end -- full -m gpu suite, smoke(), bench.py with its arguments, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_v14_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_v14_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_v14_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_v14_smoke.log
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_v14_bench_reference.json 2> gpurun_out/r02_v14_bench_reference.err
echo "reference rc=$?"
timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_v14_bench.json 2> gpurun_out/r02_v14_bench.err
echo "bench rc=$?"; grep "^\[bench\|fvgp_b200\]" gpurun_out/r02_v14_bench.err | tail -22
python - <<'PY'
import json
r = json.loads([l for l in open("gpurun_out/r02_v14_bench_reference.json") if l.startswith("{")][-1])
print("reference:", r["value"], r["ms_per_step"], r["measured_seconds_by_n"], r["same_n"], (r.get("c4") or {}).get("value"))
d = json.loads([l for l in open("gpurun_out/r02_v14_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_seconds", "gpu_launches")}, d["e2e"], d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "whole_step")}, d["roofline"]["potrf"]["seconds"], d["roofline"]["potri"])
print("kfill", d["roofline_kfill"]["frac"], d["roofline_kfill"]["lower_mode_on_the_lml_path"]["frac"])
print("ab", d.get("int8_trailing_updates_ab"))
p = d.get("parity", {})
print("parity", {k: (v.get("pass"), v.get("rel", v.get("grad_max_rel", v.get("values_max_rel")))) for k, v in p.items() if isinstance(v, dict)}, p.get("all_pass"))
print("c4", d["c4"]["value"], d["c4"]["ms_per_step"], "fresh", (d.get("c4_fresh_process") or {}).get("value"), "c1", d["c1"]["value"], "same_n", d["same_n"])
print("ratio e2e", d["e2e"]["value"] / r["value"], "same-N ratio", r["same_n"]["seconds"] / d["same_n"]["seconds"] if r["same_n"]["seconds"] else None)
PY
