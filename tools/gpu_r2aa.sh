#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py tests/test_model_gpu.py -q --tb=short -x 2>&1 | tail -15) > gpurun_out/r2aa_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2aa_pytest.log | head
timeout 120 python tools/gemm_trace.py qkv 1 2>&1 | head -12
timeout 300 python tools/gemm_sweep.py --modes 1 > gpurun_out/r2aa_sweep.md 2>&1; cat gpurun_out/r2aa_sweep.md
