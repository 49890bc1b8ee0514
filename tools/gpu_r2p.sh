#!/bin/bash
# round 2, GPU call P (1 GPU): switch-off matrix of the GEMM pipeline on the current build (CTA-pair engine, dense C2 extents)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
out=gpurun_out/r2p_switchoff.log
: > $out
for shape in qkv fc_proj; do
  for dbg in 0 1 2 4 8 9 10 3 11 13 77; do
    CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py $shape 20 2>&1 | tail -1 >> $out
  done
done
for dbg in 0 1 2 8 9 10; do
  CAPDEC_GEMM_MODE=3 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1 >> $out
done
cat $out
