#!/bin/bash
# round 2, GPU call J: lean GEMM CTAs (168 registers, > 1 KB shared memory left) -> does the AdamW overlap pay now?  + raw MMA issue rate
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r2j_pytest.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2j_pytest.log | head -20
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for sched in static dynamic; do for ov in 0 1; do
  CAPDEC_GEMM_SCHED=$sched CAPDEC_OPT_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench_${sched}_opt$ov.log 2>&1
  echo "sched=$sched opt_overlap=$ov: $(grep '"metric"' gpurun_out/r2j_bench_${sched}_opt$ov.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["roofline"]["achieved"]), round(j["full_length_captions"]["ms_per_step"],2), j["clocks"]["sm_mhz"])')"
done; done
CAPDEC_GEMM_SCHED=dynamic CAPDEC_BWD_STREAMS=0 CAPDEC_OPT_OVERLAP=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench_dynamic_1s_opt1.log 2>&1
echo "dynamic serial-bwd opt1: $(grep '"metric"' gpurun_out/r2j_bench_dynamic_1s_opt1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3))')"
echo "== raw MMA issue (no stage handshake, no loads, no epilogue) vs full, CTA pairs"
for shape in qkv fc fc_proj; do
  for dbg in 0 13 77; do
    CAPDEC_GEMM_MODE=1 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py $shape 20 2>&1 | tail -1
  done
done
for dbg in 0 13 77; do CAPDEC_GEMM_MODE=0 CAPDEC_GEMM_DBG=$dbg timeout 120 python tools/gemm_probe.py qkv 20 2>&1 | tail -1; done
