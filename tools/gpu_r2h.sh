#!/bin/bash
# round 2, GPU call H (2 GPUs): data-parallel step - plain all-reduce vs segmented per-bucket all-reduce + AdamW, x tile order
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -30) > gpurun_out/r2h_dp_pytest.log 2>&1
tail -6 gpurun_out/r2h_dp_pytest.log
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
run() { # name sched bwd_streams dp_overlap
  CAPDEC_GEMM_SCHED=$2 CAPDEC_BWD_STREAMS=$3 CAPDEC_DP_OVERLAP=$4 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 295$5 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2h_$1.log 2>&1
  echo "$1: rc=$? $(grep '"metric"' gpurun_out/r2h_$1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
}
run static_2s_plain static 1 0 11
run dynamic_1s_seg dynamic 0 2 12
run static_2s_seg static 1 2 13
run dynamic_2s_seg dynamic 1 2 14
run static_1s_seg static 0 2 15
run dynamic_1s_plain dynamic 0 0 16
