#!/bin/bash
# round 2, GPU call A: full GPU test suite on the in-pipeline 3xTF32 build, 3xTF32 GEMM probe, first fp32-grade bench line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_scale_parity_gpu.py 2>&1 | tail -40) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
(time timeout 600 python -m pytest tests/test_scale_parity_gpu.py -q 2>&1 | tail -40) > gpurun_out/r2a_scale.log 2>&1
tail -12 gpurun_out/r2a_scale.log
timeout 300 python tools/x3_probe.py > gpurun_out/r2a_x3_shapes.md 2>&1
cat gpurun_out/r2a_x3_shapes.md | tail -50
timeout 300 python tools/x3_probe.py ksweep > gpurun_out/r2a_x3_ksweep.md 2>&1
tail -40 gpurun_out/r2a_x3_ksweep.md
CAPDEC_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 > gpurun_out/r2a_bench_x3.log 2>&1
tail -1 gpurun_out/r2a_bench_x3.log | cut -c1-600
CAPDEC_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_tf32.log 2>&1
tail -1 gpurun_out/r2a_bench_tf32.log | cut -c1-600
cat gpurun_out/parity_report.jsonl | tail -60
