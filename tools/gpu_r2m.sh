#!/bin/bash
# round 2, GPU call M: full GPU suite on the final build, evidence for profiles/ (launch list, ncu --set full captures, kernel
# timeline), and the bench lines of every BASELINE config
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60) > gpurun_out/r2m_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2m_pytest.log | head -30
cp gpurun_out/parity_report.jsonl gpurun_out/r2m_parity_report.jsonl
# ---- bench lines ----
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2m_bench_c2.log 2>&1; tail -1 gpurun_out/r2m_bench_c2.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r2m_bench_reference.log 2>&1; tail -1 gpurun_out/r2m_bench_reference.log | cut -c1-700
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for w in c1 c3 c4 c5; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2m_bench_$w.log 2>&1
  echo "$w: $(grep '"metric"' gpurun_out/r2m_bench_$w.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]))')"
done
timeout 400 python bench.py --workload c5 --precision tf32x3 --steps 5 --warmup 3 > gpurun_out/r2m_bench_c5_x3.log 2>&1
echo "c5 x3: $(grep '"metric"' gpurun_out/r2m_bench_c5_x3.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]))')"
timeout 300 python bench.py --full_length --steps 20 --warmup 5 > gpurun_out/r2m_bench_full_length.log 2>&1
echo "full_length: $(grep '"metric"' gpurun_out/r2m_bench_full_length.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]))')"
# ---- evidence ----
timeout 300 python tools/step_timeline.py > gpurun_out/r2m_step_timeline.md 2>gpurun_out/r2m_step_timeline.err; head -20 gpurun_out/r2m_step_timeline.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2m_launches_step.csv python tools/profile_step.py --steps 1 > gpurun_out/r2m_profile_step.log 2>&1
python tools/profile_step.py --summarise gpurun_out/r2m_launches_step.csv > gpurun_out/r2m_launches_step.md; cat gpurun_out/r2m_launches_step.md | head -16
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 6 -c 1 -f -o gpurun_out/r2m_ncu_qkv_gemm python tools/gemm_probe.py qkv 10 > /dev/null 2>&1
ncu -i gpurun_out/r2m_ncu_qkv_gemm.ncu-rep --page raw --csv > gpurun_out/r2m_ncu_qkv_gemm.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc_bwd8 -s 3 -c 1 -f -o gpurun_out/r2m_ncu_attn_bwd8 python tools/profile_step.py --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/r2m_ncu_attn_bwd8.ncu-rep --page raw --csv > gpurun_out/r2m_ncu_attn_bwd8.csv 2>/dev/null
CAPDEC_X3=1 timeout 300 ncu --set full --clock-control none -k regex:gemm_tf32 -s 4 -c 1 -f -o gpurun_out/r2m_ncu_qkv_gemm_x3 python tools/x3_probe.py qkv > /dev/null 2>&1
ncu -i gpurun_out/r2m_ncu_qkv_gemm_x3.ncu-rep --page raw --csv > gpurun_out/r2m_ncu_qkv_gemm_x3.csv 2>/dev/null
timeout 300 python tools/x3_probe.py > gpurun_out/r2m_x3_shapes.md 2>&1
timeout 300 python tools/cublas_compare.py > gpurun_out/r2m_cublas_compare.md 2>&1; cat gpurun_out/r2m_cublas_compare.md
ls -la gpurun_out | tail -30
