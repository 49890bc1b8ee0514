set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -15) > gpurun_out/s14_pytest.log 2>&1
tail -4 gpurun_out/s14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s14_smoke.log 2>&1
tail -2 gpurun_out/s14_smoke.log
for w in c1 c3 c4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 4 > gpurun_out/s14_bench_$w.log 2>&1; tail -1 gpurun_out/s14_bench_$w.log | cut -c1-160; done
timeout 300 python bench.py > gpurun_out/s14_bench.log 2>&1
tail -1 gpurun_out/s14_bench.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s14_bench_ref.log 2>&1
tail -1 gpurun_out/s14_bench_ref.log | cut -c1-300
