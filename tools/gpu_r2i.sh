#!/bin/bash
# round 2, GPU call I (N GPUs): data-parallel tests, then the N-GPU bench with the sharded optimizer (default) and with all-reduce + full update
cd ${GRAFT_REPO_ROOT:-.}
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  (time timeout 600 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -30) > gpurun_out/r2i_dp_pytest.log 2>&1
  tail -8 gpurun_out/r2i_dp_pytest.log
fi
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
run() { # name sharded port
  NCCL_DEBUG=${NCCL_DEBUG:-WARN} CAPDEC_DP_SHARDED=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 296$3 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2i_n${N}_$1.log 2>&1
  echo "N=$N $1: rc=$? $(grep '"metric"' gpurun_out/r2i_n${N}_$1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
}
run sharded 1 21
run allreduce 0 22
run sharded_b 1 23
