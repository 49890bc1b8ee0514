#!/bin/bash
# round 2, GPU call O (1 GPU): GEMM A-B sweeps - tile raster order, TMA L2 promotion, pipeline depth
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() { # name env...
  n=$1; shift
  env "$@" timeout 300 python tools/gemm_sweep.py > gpurun_out/r2o_sweep_$n.md 2>&1
  echo "== $n"; cat gpurun_out/r2o_sweep_$n.md
}
run base CAPDEC_X=0
run raster1 CAPDEC_GEMM_RASTER=1
run promo128 CAPDEC_GEMM_L2PROMO=2
run promo0 CAPDEC_GEMM_L2PROMO=0
run stages4 CAPDEC_GEMM_STAGES=4
