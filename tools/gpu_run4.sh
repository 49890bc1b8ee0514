set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40) > gpurun_out/s6_pytest.log 2>&1
tail -4 gpurun_out/s6_pytest.log
for m in 1 2 3; do for s in qkv fc_proj; do CAPDEC_GEMM_DBG=10 CAPDEC_GEMM_MODE=$m timeout 120 python tools/gemm_probe.py $s 20; done; done > gpurun_out/s6_loadsonly.log 2>&1
cat gpurun_out/s6_loadsonly.log
timeout 300 python tools/op_probe.py gemm fcproj_dgrad_mul > gpurun_out/s6_gemm.log 2>&1
cat gpurun_out/s6_gemm.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench.log 2>&1
tail -1 gpurun_out/s6_bench.log | cut -c1-300
