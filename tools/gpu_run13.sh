set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_packed_gpu.py tests/test_model_gpu.py tests/test_datafeed_gpu.py -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -15) > gpurun_out/s15_pytest.log 2>&1
tail -3 gpurun_out/s15_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s15_bench_optov.log 2>&1
tail -1 gpurun_out/s15_bench_optov.log | cut -c1-200
CAPDEC_OPT_OVERLAP=0 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s15_bench_noov.log 2>&1
tail -1 gpurun_out/s15_bench_noov.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s15_bench_optov2.log 2>&1
tail -1 gpurun_out/s15_bench_optov2.log | cut -c1-200
