#!/bin/bash
# First GPU call of the next round: run what was written after the round-1 GPU budget was spent (all of it is opt-in and
# outside every number in DESIGN.md until this passes).
#   gpurun --timeout 900 -- 'bash tools/gpu_validate_experimental.sh'                 (1 GPU: the fit loop)
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_validate_experimental.sh dp'    (2 GPUs: segmented-graph DP overlap)
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
export CAPDEC_TEST_EXPERIMENTAL=1
if [ "$1" == "dp" ]; then
  (time timeout 600 python -m pytest tests/test_dp_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/exp_dp_pytest.log 2>&1
  tail -4 gpurun_out/exp_dp_pytest.log
  for ov in 0 2; do
    CAPDEC_DP_OVERLAP=$ov CAPDEC_BENCH_NO_CPU=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/exp_bench_dp_overlap$ov.log 2>&1
    tail -1 gpurun_out/exp_bench_dp_overlap$ov.log | cut -c1-300
  done
else
  (time timeout 600 python -m pytest tests/test_fit_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/exp_fit_pytest.log 2>&1
  tail -4 gpurun_out/exp_fit_pytest.log
fi
