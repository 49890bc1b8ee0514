set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -15) > gpurun_out/s4_pytest_gemm.log 2>&1
tail -4 gpurun_out/s4_pytest_gemm.log
timeout 300 python tools/op_probe.py gemm fcproj_dgrad_mul fc_fwd lm_dgrad > gpurun_out/s4_gemm.log 2>&1
cat gpurun_out/s4_gemm.log
CAPDEC_GEMM_TUNE_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench.log 2> gpurun_out/s4_tune.log
tail -1 gpurun_out/s4_bench.log | cut -c1-400
CAPDEC_GEMM_AUTOTUNE=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench_notune.log 2>&1
tail -1 gpurun_out/s4_bench_notune.log | cut -c1-400
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/s4_pytest_all.log 2>&1
tail -4 gpurun_out/s4_pytest_all.log
