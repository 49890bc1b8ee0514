#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
out=gpurun_out/r2x_trace.md
: > $out
timeout 120 python tools/gemm_trace.py attn_proj 1 8820 >> $out 2>&1
timeout 120 python tools/gemm_trace.py qkv 1 >> $out 2>&1
cat $out
