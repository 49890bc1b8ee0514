set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py -x -q -m gpu 2>&1 | tail -25) > gpurun_out/s3_pytest_ops.log 2>&1
tail -5 gpurun_out/s3_pytest_ops.log
timeout 300 python tools/op_probe.py ln > gpurun_out/s3_ln.log 2>&1
CAPDEC_LN_BWD_PIPE=0 timeout 300 python tools/op_probe.py ln > gpurun_out/s3_ln_old.log 2>&1
timeout 300 python tools/op_probe.py attn > gpurun_out/s3_attn.log 2>&1
CAPDEC_ATTN_BWD8=0 timeout 300 python tools/op_probe.py attn > gpurun_out/s3_attn_old.log 2>&1
timeout 600 python tools/op_probe.py gemm > gpurun_out/s3_gemm.log 2>&1
cat gpurun_out/s3_ln.log gpurun_out/s3_ln_old.log gpurun_out/s3_attn.log gpurun_out/s3_attn_old.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.log 2>&1
tail -2 gpurun_out/s3_bench.log | cut -c1-600
