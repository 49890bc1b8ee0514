set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 150 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/s12_dp_pytest.log 2>&1
tail -3 gpurun_out/s12_dp_pytest.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s12_bench2_pipe.log 2>&1
tail -1 gpurun_out/s12_bench2_pipe.log | cut -c1-200
CAPDEC_DP_PIPELINE=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s12_bench2_plain.log 2>&1
tail -1 gpurun_out/s12_bench2_plain.log | cut -c1-200
