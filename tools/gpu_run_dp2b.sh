set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CAPDEC_DP_OVERLAP=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s11_bench2_overlap.log 2>&1
tail -1 gpurun_out/s11_bench2_overlap.log | cut -c1-200
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s11_bench2.log 2>&1
tail -1 gpurun_out/s11_bench2.log | cut -c1-200
