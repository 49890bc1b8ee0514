#!/bin/bash
# round 2, GPU call R (1 GPU): what the epilogue of one tile costs, piece by piece (no loads, no MMA: barrier skeleton + epilogue)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
out=gpurun_out/r2r_epilogue_pieces.log
: > $out
for dbg in 11 3 7 23 55 183 131 0 4 20 52 180 8; do
  CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1 >> $out
done
cat $out
