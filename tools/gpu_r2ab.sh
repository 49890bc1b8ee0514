#!/bin/bash
# round 2, GPU call AB (N GPUs): the N-GPU bench with the peer-memory update (default) and with NCCL reduce-scatter / all-gather
cd ${GRAFT_REPO_ROOT:-.}
N=${1:-8}
mkdir -p gpurun_out
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
run() { # name peer port
  env NCCL_DEBUG=${NCCL_DEBUG:-WARN} CAPDEC_DP_PEER=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 298$3 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2ab_n${N}_$1.log 2>&1
  echo "N=$N $1: rc=$? $(grep '"metric"' gpurun_out/r2ab_n${N}_$1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["config"].get("dp_update","")[:40])')"
  grep -i "unavailable\|error" gpurun_out/r2ab_n${N}_$1.log | head -3
}
run peer 1 41
run nccl 0 42
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab_n1.log 2>&1
echo "N=1: $(grep '"metric"' gpurun_out/r2ab_n1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
