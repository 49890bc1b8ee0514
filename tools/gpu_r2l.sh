#!/bin/bash
# round 2, GPU call L: half-width tail tiles: GEMM tests, parity subset, bench A/B, single-GEMM probe A/B
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(time timeout 900 python -m pytest tests/test_ops_gpu.py -q --tb=short -x -k "gemm" 2>&1 | tail -30) > gpurun_out/r2l_pytest_gemm.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2l_pytest_gemm.log | head -20
(time timeout 900 python -m pytest tests/test_model_gpu.py tests/test_scale_parity_gpu.py tests/test_packed_gpu.py tests/test_decode_gpu.py -q --tb=short 2>&1 | tail -30) > gpurun_out/r2l_pytest_model.log 2>&1
grep -E "^E  |FAILED|passed|failed" gpurun_out/r2l_pytest_model.log | head -20
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
for t in 1 0 1 0; do
  CAPDEC_GEMM_TAIL=$t timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2l_bench_tail$t.log 2>&1
  echo "tail=$t: $(grep '"metric"' gpurun_out/r2l_bench_tail$t.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), round(j["roofline"]["achieved"]), round(j["full_length_captions"]["ms_per_step"],2), j["clocks"]["sm_mhz"])')"
done
for t in 1 0; do echo "== tail=$t"; CAPDEC_GEMM_TAIL=$t timeout 300 python tools/op_probe.py gemm qkv_fwd aproj_fwd fc_fwd fcproj_fwd qkv_dgrad fc_dgrad fcproj_dgrad_mul 2>&1 | cut -c1-200; done
