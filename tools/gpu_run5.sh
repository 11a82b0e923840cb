#!/bin/bash
# GPU run 5 (round 2, 1 GPU): the N=1 bench line with every sub-record, Ozaki POTRF A/B + parity, POTRS / POTRF probe
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_v5_bench.json 2> gpurun_out/r02_v5_bench.err
echo "bench rc=$?"; grep "^\[bench\|fvgp_b200\]" gpurun_out/r02_v5_bench.err | tail -30
for OZ in 0 8; do
FVGP_OZAKI=$OZ timeout 300 python tools/potrf_sweep.py 50000 > gpurun_out/r02_v5_potrf_50k_ozaki$OZ.log 2>&1
echo "potrf 50k ozaki=$OZ rc=$?"; tail -3 gpurun_out/r02_v5_potrf_50k_ozaki$OZ.log
done
FVGP_OZAKI=8 timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -k "benchmarked_n or 16000 or 8000" > gpurun_out/r02_v5_pytest_ozaki_parity.log 2>&1
echo "ozaki parity rc=$?"; tail -8 gpurun_out/r02_v5_pytest_ozaki_parity.log
timeout 600 python tools/potrf_probe.py > gpurun_out/r02_v5_potrf_probe.log 2>&1
echo "potrf probe rc=$?"; tail -9 gpurun_out/r02_v5_potrf_probe.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_at_size.py::test_c4_fifty_blocks_of_the_1m_pattern_are_bit_exact > gpurun_out/r02_v5_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_v5_pytest_all.log
