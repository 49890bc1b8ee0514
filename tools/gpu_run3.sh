set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_packed_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/s5_pytest.log 2>&1
tail -4 gpurun_out/s5_pytest.log
CAPDEC_GEMM_TUNE_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s5_bench.log 2> gpurun_out/s5_tune.log
tail -1 gpurun_out/s5_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s5_launches.csv python tools/profile_step.py --steps 1 > gpurun_out/s5_prof.log 2>&1
tail -2 gpurun_out/s5_prof.log
