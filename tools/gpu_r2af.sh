#!/bin/bash
# round 2, GPU call AF (8 GPUs): the 8-GPU bench with the push variant of the peer-memory update (default), N = 1 on the same box
cd ${GRAFT_REPO_ROOT:-.}
N=${1:-8}
mkdir -p gpurun_out
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
env NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29861 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2af_n${N}_push.log 2>&1
echo "N=$N push: rc=$? $(grep '"metric"' gpurun_out/r2af_n${N}_push.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["config"].get("dp_update","")[:50])')"
grep -i "unavailable\|error" gpurun_out/r2af_n${N}_push.log | head -3
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2af_n1.log 2>&1
echo "N=1: $(grep '"metric"' gpurun_out/r2af_n1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
