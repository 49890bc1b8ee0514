#!/bin/bash
# round 2, GPU call U (1 GPU): the epilogue under load, piece by piece, now that the mainloop runs at the MMA rate
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
out=gpurun_out/r2u_epilogue_under_load.log
: > $out
for dbg in 0 4 20 52 180 132 8; do
  CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1 >> $out
done
for d in 2 3 4; do
  echo -n "cdepth=$d " >> $out
  CAPDEC_GEMM_CDEPTH=$d CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1 >> $out
done
cat $out
