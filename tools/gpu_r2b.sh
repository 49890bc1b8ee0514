#!/bin/bash
# round 2, GPU call B: full GPU suite (softmax-backward row sum in the attention kernels), C3 bisection of the mapper
# query-weight gradient (exact FFMA attention vs tensor-core attention), x3 numbers again
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2b_pytest.log 2>&1
tail -25 gpurun_out/r2b_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_report.jsonl
rm -f gpurun_out/parity_report.jsonl
CAPDEC_ATTN_IMPL=ffma CAPDEC_PACKED=0 timeout 600 python -m pytest tests/test_scale_parity_gpu.py -q -k "c3" 2>&1 | tail -5
mv gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_c3_ffma_attention.jsonl
cat gpurun_out/r2b_parity_c3_ffma_attention.jsonl
grep scale_parity gpurun_out/r2b_parity_report.jsonl
