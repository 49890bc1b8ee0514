#!/bin/bash
# round 2, GPU call F: dynamic order with the cheap slot release: full suite (verbose failures), bench dynamic/static x AdamW overlap, x3 bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150) > gpurun_out/r2f_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2f_pytest.log | head -60
cp gpurun_out/parity_report.jsonl gpurun_out/r2f_parity_report.jsonl
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for sched in dynamic static; do for ov in 0 1; do
  CAPDEC_GEMM_SCHED=$sched CAPDEC_OPT_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2f_bench_${sched}_opt$ov.log 2>&1
  echo "sched=$sched opt_overlap=$ov: $(grep '"metric"' gpurun_out/r2f_bench_${sched}_opt$ov.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["achieved"], j["full_length_captions"]["ms_per_step"], j["clocks"])')"
done; done
CAPDEC_OPT_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 > gpurun_out/r2f_bench_x3.log 2>&1
grep '"metric"' gpurun_out/r2f_bench_x3.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("x3", j["ms_per_step"], j["value"])'
timeout 200 python tools/x3_probe.py qkv fc fc_proj qkv_wgrad > gpurun_out/r2f_x3_shapes.md 2>&1; grep tf32x3 gpurun_out/r2f_x3_shapes.md
