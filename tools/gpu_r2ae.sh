#!/bin/bash
# round 2, GPU call AE (1 GPU): evidence of the final round-2 build - bench lines of every BASELINE config, kernel timeline,
# ncu launch list, ncu --set full captures of the roofline anchor and of the attention backward, cuBLAS comparison
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2ae_bench_c2.log 2>&1; tail -1 gpurun_out/r2ae_bench_c2.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ae_bench_reference.log 2>&1; tail -1 gpurun_out/r2ae_bench_reference.log | cut -c1-300
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for w in c1 c3 c4; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2ae_bench_$w.log 2>&1
  echo "$w: $(grep '"metric"' gpurun_out/r2ae_bench_$w.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]))')"
done
timeout 300 python bench.py --full_length --steps 20 --warmup 5 > gpurun_out/r2ae_bench_full_length.log 2>&1
timeout 300 python tools/step_timeline.py > gpurun_out/r2ae_step_timeline.md 2>gpurun_out/r2ae_step_timeline.err; head -24 gpurun_out/r2ae_step_timeline.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ae_launches_step.csv python tools/profile_step.py --steps 1 > gpurun_out/r2ae_profile_step.log 2>&1
python tools/profile_step.py --summarise gpurun_out/r2ae_launches_step.csv > gpurun_out/r2ae_launches_step.md; cat gpurun_out/r2ae_launches_step.md | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 6 -c 1 -f -o gpurun_out/r2ae_ncu_qkv_gemm python tools/gemm_probe.py qkv 10 > /dev/null 2>&1
ncu -i gpurun_out/r2ae_ncu_qkv_gemm.ncu-rep --page raw --csv > gpurun_out/r2ae_ncu_qkv_gemm.csv 2>/dev/null
timeout 300 python tools/cublas_compare.py > gpurun_out/r2ae_cublas_compare.md 2>&1; cat gpurun_out/r2ae_cublas_compare.md
rm -f gpurun_out/r2ae_ncu_qkv_gemm.ncu-rep
