#!/bin/bash
# round 2, GPU call Q (1 GPU): epilogue rework (TMEM loads one chunk ahead, staging 2..4 deep) - correctness, then A-B timings
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py -q --tb=short -x -k "gemm or packed or adamw" 2>&1 | tail -15) > gpurun_out/r2q_pytest.log 2>&1
tail -6 gpurun_out/r2q_pytest.log
out=gpurun_out/r2q_epilogue.log
: > $out
for depth in 2 3 4; do
  for shape in qkv fc_proj attn_proj; do
    for dbg in 0 4 8; do
      echo -n "cdepth=$depth " >> $out
      CAPDEC_GEMM_CDEPTH=$depth CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py $shape 20 2>&1 | tail -1 >> $out
    done
  done
done
echo -n "auto fc act4 aux " >> $out; CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py fc 20 4 1 2>&1 | tail -1 >> $out
echo -n "cdepth=2 fc act4 aux " >> $out; CAPDEC_GEMM_CDEPTH=2 CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py fc 20 4 1 2>&1 | tail -1 >> $out
cat $out
timeout 300 python tools/gemm_sweep.py > gpurun_out/r2q_sweep.md 2>&1; cat gpurun_out/r2q_sweep.md
