#!/bin/bash
# round 2, GPU call AI (1 GPU): tile raster for the LM head (B = wte does not fit L2), then the full GPU suite + bench on the final build
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for r in 0 1; do
  echo -n "raster=$r "; CAPDEC_GEMM_RASTER=$r CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py lm_head 10 2>&1 | tail -1
done
echo -n "auto "; CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py lm_head 10 2>&1 | tail -1
(time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2ai_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2ai_pytest.log | head -20
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for r in auto 0; do
  if [ "$r" == "0" ]; then export CAPDEC_GEMM_RASTER=0; fi
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2ai_bench_raster_$r.log 2>&1
  grep '"metric"' gpurun_out/r2ai_bench_raster_$r.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("raster '$r':", round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], round(j["roofline"]["frac"],3))'
done
