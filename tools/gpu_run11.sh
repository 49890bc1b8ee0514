set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(CAPDEC_BWD_STREAMS=1 timeout 600 python -m pytest tests/test_packed_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -15) > gpurun_out/s13_pytest_streams.log 2>&1
tail -3 gpurun_out/s13_pytest_streams.log
CAPDEC_BWD_STREAMS=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s13_bench_streams.log 2>&1
tail -1 gpurun_out/s13_bench_streams.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s13_bench_serial.log 2>&1
tail -1 gpurun_out/s13_bench_serial.log | cut -c1-200
CAPDEC_BWD_STREAMS=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s13_bench_streams2.log 2>&1
tail -1 gpurun_out/s13_bench_streams2.log | cut -c1-200
