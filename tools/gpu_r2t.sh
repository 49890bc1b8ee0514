#!/bin/bash
# round 2, GPU call T (1 GPU): converged-warp producer / MMA issuer - correctness, CTA timeline, sweep
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py -q --tb=short -x -k "gemm or packed or adamw" 2>&1 | tail -15) > gpurun_out/r2t_pytest.log 2>&1
tail -4 gpurun_out/r2t_pytest.log
out=gpurun_out/r2t_trace.md
: > $out
timeout 120 python tools/gemm_trace.py qkv 1 >> $out 2>&1
timeout 120 python tools/gemm_trace.py attn_proj 1 8820 >> $out 2>&1
cat $out
timeout 300 python tools/gemm_sweep.py > gpurun_out/r2t_sweep.md 2>&1; cat gpurun_out/r2t_sweep.md
for dbg in 0 8 10; do CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1; done
