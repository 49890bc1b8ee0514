#!/bin/bash
# round 2, GPU call E: where does the dynamic tile order lose time?  single-GEMM sweep under four scheduler variants
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
SH="qkv_fwd fcproj_fwd fc_dgrad qkv_wgrad fcproj_dgrad_mul lm_head"
for v in "static:0" "dynamic:0" "dynamic:16" "dynamic:32" "dynamic:48"; do
  s=${v%%:*}; d=${v##*:}
  echo "=== sched=$s dbg=$d"
  CAPDEC_GEMM_SCHED=$s CAPDEC_GEMM_DBG=$d timeout 300 python tools/op_probe.py gemm $SH 2>&1 | tee gpurun_out/r2e_probe_${s}_$d.log | cut -c1-330
done
