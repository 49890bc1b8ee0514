#!/bin/bash
# round 2, GPU call G: where does the dynamic order lose 1.3 ms in the step?  scheduler x two-stream backward x AdamW overlap
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for sched in static dynamic; do for bs in 1 0; do for ov in 0 1; do
  CAPDEC_GEMM_SCHED=$sched CAPDEC_BWD_STREAMS=$bs CAPDEC_OPT_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2g_bench_${sched}_bs${bs}_opt$ov.log 2>&1
  echo "sched=$sched bwd_streams=$bs opt_overlap=$ov: $(grep '"metric"' gpurun_out/r2g_bench_${sched}_bs${bs}_opt$ov.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["roofline"]["achieved"]), round(j["full_length_captions"]["ms_per_step"],2), j["clocks"]["sm_mhz"])')"
done; done; done
for sched in static dynamic; do
  CAPDEC_GEMM_SCHED=$sched timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_launches_$sched.csv python tools/profile_step.py --steps 1 > gpurun_out/r2g_profile_$sched.log 2>&1
  python tools/profile_step.py --summarise gpurun_out/r2g_launches_$sched.csv > gpurun_out/r2g_launches_$sched.md
  echo "== $sched"; head -12 gpurun_out/r2g_launches_$sched.md; tail -1 gpurun_out/r2g_launches_$sched.md
done
