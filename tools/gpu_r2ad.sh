#!/bin/bash
# round 2, GPU call AD (1 GPU): programmatic dependent launch of the GEMMs - full GPU suite, then bench A-B on the same box
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2ad_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2ad_pytest.log | head -20
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for pdl in 1 0 1 0; do
  CAPDEC_GEMM_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2ad_bench_pdl$pdl.log 2>&1
  grep '"metric"' gpurun_out/r2ad_bench_pdl$pdl.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("pdl='$pdl':", round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], round(j["roofline"]["frac"],3), j.get("full_length_captions",{}).get("ms_per_step"))'
done
