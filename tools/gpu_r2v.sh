#!/bin/bash
# round 2, GPU call V (1 GPU): LSU store path of the epilogue - correctness, A-B, CTA timeline, sweep
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py -q --tb=short -x -k "gemm or packed" 2>&1 | tail -15) > gpurun_out/r2v_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2v_pytest.log
for lsu in 1 0; do for shape in qkv fc_proj attn_proj; do
  echo -n "lsu=$lsu "; CAPDEC_GEMM_LSU_STORE=$lsu CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py $shape 20 2>&1 | tail -1
done; done
echo -n "lsu=1 fc act4 aux: "; CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py fc 20 4 1 2>&1 | tail -1
echo -n "lsu=0 fc act4 aux: "; CAPDEC_GEMM_LSU_STORE=0 CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py fc 20 4 1 2>&1 | tail -1
timeout 120 python tools/gemm_trace.py qkv 1 > gpurun_out/r2v_trace.md 2>&1; cat gpurun_out/r2v_trace.md
timeout 300 python tools/gemm_sweep.py > gpurun_out/r2v_sweep.md 2>&1; cat gpurun_out/r2v_sweep.md
