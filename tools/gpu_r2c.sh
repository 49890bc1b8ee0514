#!/bin/bash
# round 2, GPU call C (2 GPUs): data-parallel tests incl. the segmented-graph overlap (CAPDEC_DP_OVERLAP=2), and the
# 2-GPU bench with and without it
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -30) > gpurun_out/r2c_dp_pytest.log 2>&1
tail -8 gpurun_out/r2c_dp_pytest.log
for ov in 0 2; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL CAPDEC_DP_OVERLAP=$ov CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2c_bench_dp2_overlap$ov.log 2>&1
  echo "overlap=$ov rc=$?"
  grep '"metric"' gpurun_out/r2c_bench_dp2_overlap$ov.log | tail -1 | cut -c1-400
  grep -E "NVLS|Ring|Tree|algo|Connected all" gpurun_out/r2c_bench_dp2_overlap$ov.log | head -5
done
CAPDEC_BENCH_NO_CPU=1 CAPDEC_BENCH_NO_X3=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2c_bench_dp1.log 2>&1
grep '"metric"' gpurun_out/r2c_bench_dp1.log | tail -1 | cut -c1-400
