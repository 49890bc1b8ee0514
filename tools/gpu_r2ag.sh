#!/bin/bash
# round 2, GPU call AG (2 GPUs): gradient clearing moved into the forward pass (side stream) - parity tests, bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -40) > gpurun_out/r2ag_dp_pytest.log 2>&1
tail -6 gpurun_out/r2ag_dp_pytest.log
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
env NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ag_n2_push.log 2>&1
echo "N=2 push: rc=$? $(grep '"metric"' gpurun_out/r2ag_n2_push.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag_n1.log 2>&1
echo "N=1: $(grep '"metric"' gpurun_out/r2ag_n1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
