#!/bin/bash
# round 2, GPU call D: dynamic tile scheduler (default) vs static order, with and without the AdamW overlap; full GPU suite;
# library comparison table and what cuBLAS launches for the fc shape
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2d_pytest.log 2>&1
tail -15 gpurun_out/r2d_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2d_parity_report.jsonl
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for sched in dynamic static; do for ov in 0 1; do
  CAPDEC_GEMM_SCHED=$sched CAPDEC_OPT_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2d_bench_${sched}_opt$ov.log 2>&1
  echo "sched=$sched opt_overlap=$ov: $(grep '"metric"' gpurun_out/r2d_bench_${sched}_opt$ov.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["achieved"], j["clocks"])')"
done; done
timeout 600 python tools/cublas_compare.py > gpurun_out/r2d_cublas_compare.md 2>&1
cat gpurun_out/r2d_cublas_compare.md
timeout 300 ncu --set full --clock-control none -k regex:'gemm|cutlass|sm100|nvjet|xmma' -s 8 -c 2 -o gpurun_out/r2d_ncu_cublas_fc python tools/cublas_compare.py lib fc > gpurun_out/r2d_ncu_cublas_fc.log 2>&1
tail -3 gpurun_out/r2d_ncu_cublas_fc.log
ncu -i gpurun_out/r2d_ncu_cublas_fc.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_cublas_fc.csv 2>/dev/null
ls -la gpurun_out | tail -12
