#!/bin/bash
# round 2, GPU call K: split-K (zero + reduce-add) for the N = 768 linear GEMMs: parity subset, then bench A/B
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 900 python -m pytest tests/test_model_gpu.py tests/test_scale_parity_gpu.py tests/test_packed_gpu.py tests/test_fit_gpu.py -q --tb=short 2>&1 | tail -30) > gpurun_out/r2k_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2k_pytest.log | head -20
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for sk in 1 0 1 0; do
  CAPDEC_SPLITK_LINEAR=$sk CAPDEC_GEMM_TUNE_VERBOSE=$sk timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2k_bench_sk$sk.log 2>&1
  echo "splitk_linear=$sk: $(grep '"metric"' gpurun_out/r2k_bench_sk$sk.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), round(j["roofline"]["achieved"]), round(j["full_length_captions"]["ms_per_step"],2), j["clocks"]["sm_mhz"])')"
done
grep "capdec gemm tune" gpurun_out/r2k_bench_sk1.log | grep "acc=1" | sort | uniq -c | sort -rn | head -30
CAPDEC_SPLITK_LINEAR=1 timeout 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 > gpurun_out/r2k_bench_x3.log 2>&1
grep '"metric"' gpurun_out/r2k_bench_x3.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("x3", j["ms_per_step"], j["value"])'
