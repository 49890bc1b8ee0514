set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/s8_dp_pytest.log 2>&1
tail -3 gpurun_out/s8_dp_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s8_bench2.log 2>&1
tail -1 gpurun_out/s8_bench2.log | cut -c1-300
CAPDEC_DP_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s8_bench2_overlap.log 2>&1
tail -1 gpurun_out/s8_bench2_overlap.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s8_bench2_ref.log 2>&1
tail -1 gpurun_out/s8_bench2_ref.log | cut -c1-300
