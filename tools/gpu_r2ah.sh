#!/bin/bash
# round 2, GPU call AH (1 GPU): staging depth / stage count A-B on the final pipeline, then what the driver runs: smoke(), default bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for d in 2 3 4; do for shape in qkv fc_proj attn_proj; do
  echo -n "cdepth=$d "; CAPDEC_GEMM_CDEPTH=$d CAPDEC_GEMM_MODE=1 timeout 120 python tools/gemm_probe.py $shape 30 2>&1 | tail -1
done; done
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2ah_bench_default.log 2>&1; tail -1 gpurun_out/r2ah_bench_default.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("default bench:", j["steps"], j["warmup"], round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"], round(j["roofline"]["frac"],3), "x3:", j.get("fp32_grade",{}).get("ms_per_step"), "c5:", j.get("c5",{}).get("value"), "cpu:", j.get("cpu_baseline",{}).get("value"))'
