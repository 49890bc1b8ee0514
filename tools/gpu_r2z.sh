#!/bin/bash
# round 2, GPU call Z (1 GPU): kEpi instantiations: full GPU suite, probes of the fused epilogues, bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2z_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2z_pytest.log | head -20
echo -n "fc act4 aux: "; CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py fc 20 4 1 2>&1 | tail -1
timeout 200 python tools/op_probe.py gemm 2>&1 | tail -25
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2z_bench_c2.log 2>&1
grep '"metric"' gpurun_out/r2z_bench_c2.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("c2:", round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], round(j["roofline"]["frac"],3), j.get("full_length_captions",{}).get("ms_per_step"))'
