#!/bin/bash
# round 2, GPU call AC (N GPUs): copy-engine gradient pushes - parity tests (N=2), then bench push / pull / NCCL
cd ${GRAFT_REPO_ROOT:-.}
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  (time timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -40) > gpurun_out/r2ac_dp_pytest.log 2>&1
  tail -12 gpurun_out/r2ac_dp_pytest.log
fi
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
run() { # name port env...
  n=$1; port=$2; shift; shift
  env NCCL_DEBUG=${NCCL_DEBUG:-WARN} "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2ac_n${N}_$n.log 2>&1
  echo "N=$N $n: rc=$? $(grep '"metric"' gpurun_out/r2ac_n${N}_$n.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["config"].get("dp_update","")[:60])')"
  grep -i "unavailable\|error" gpurun_out/r2ac_n${N}_$n.log | head -3
}
run push 29851 CAPDEC_X=0
run pull 29852 CAPDEC_DP_PUSH=0
run nccl 29853 CAPDEC_DP_PEER=0
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ac_n1.log 2>&1
echo "N=1: $(grep '"metric"' gpurun_out/r2ac_n1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
