#!/bin/bash
# round 2, GPU call W (1 GPU): converged-warp GEMM roles + elected epilogue stores + reciprocal tile decode: full GPU suite, sweep, timeline, bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2w_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2w_pytest.log | head -20
timeout 120 python tools/gemm_trace.py qkv 1 > gpurun_out/r2w_trace.md 2>&1; cat gpurun_out/r2w_trace.md
timeout 300 python tools/gemm_sweep.py > gpurun_out/r2w_sweep.md 2>&1; cat gpurun_out/r2w_sweep.md
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2w_bench_c2.log 2>&1
grep '"metric"' gpurun_out/r2w_bench_c2.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("c2:", round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], round(j["roofline"]["frac"],3), j.get("full_length_captions",{}).get("ms_per_step"))'
