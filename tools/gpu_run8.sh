set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -30) > gpurun_out/s10_pytest.log 2>&1
tail -4 gpurun_out/s10_pytest.log
timeout 200 python tools/op_probe.py mul > gpurun_out/s10_mul.log 2>&1
timeout 200 python tools/op_probe.py ln > gpurun_out/s10_ln.log 2>&1
cat gpurun_out/s10_mul.log gpurun_out/s10_ln.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/s10_bench.log 2>&1
tail -1 gpurun_out/s10_bench.log | cut -c1-300
