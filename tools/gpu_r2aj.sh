#!/bin/bash
# round 2, GPU call AJ (2 GPUs): the driver's multi-GPU launch on the final build (sanity)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29881 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2aj_n2.log 2>&1
echo "N=2: rc=$? $(grep '"metric"' gpurun_out/r2aj_n2.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["config"].get("dp_update","")[:40])')"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29882 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-200
