#!/bin/bash
# round 2, GPU call N (N GPUs): the peer-memory optimizer step (csrc/peer.cu) - kernel test on one GPU, 2-rank parity tests, then
# the N-GPU bench with the peer kernel (default) and with NCCL reduce-scatter / all-gather (CAPDEC_DP_PEER=0)
cd ${GRAFT_REPO_ROOT:-.}
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  (time timeout 300 python -m pytest tests/test_ops_gpu.py -q --tb=short -k "adamw" 2>&1 | tail -15) > gpurun_out/r2n_peer_kernel_pytest.log 2>&1
  tail -4 gpurun_out/r2n_peer_kernel_pytest.log
  (time timeout 600 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -40) > gpurun_out/r2n_dp_pytest.log 2>&1
  tail -12 gpurun_out/r2n_dp_pytest.log
fi
export CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1
run() { # name peer port [extra env]
  env NCCL_DEBUG=${NCCL_DEBUG:-WARN} CAPDEC_DP_PEER=$2 $4 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 297$3 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2n_n${N}_$1.log 2>&1
  echo "N=$N $1: rc=$? $(grep '"metric"' gpurun_out/r2n_n${N}_$1.log | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["config"].get("dp_update","")[:40])')"
  grep -i "unavailable\|error" gpurun_out/r2n_n${N}_$1.log | head -3
}
run peer 1 31
run nccl 0 32
run peer_b 1 33
run peer_syncloss 1 34 CAPDEC_BENCH_SYNC_LOSS=1
