set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -40) > gpurun_out/s9_pytest.log 2>&1
tail -4 gpurun_out/s9_pytest.log
timeout 200 python tools/op_probe.py mul > gpurun_out/s9_mul.log 2>&1
CAPDEC_GEMM_DBG=4 timeout 200 python tools/op_probe.py mul > gpurun_out/s9_mul_nostore.log 2>&1
cat gpurun_out/s9_mul.log gpurun_out/s9_mul_nostore.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s9_bench.log 2>&1
tail -1 gpurun_out/s9_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s9_launches.csv python tools/profile_step.py --steps 1 > gpurun_out/s9_prof.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:add_ln_bwd_pipe -s 3 -c 1 -f -o gpurun_out/r1_ncu_ln_bwd_pipe python tools/profile_step.py --steps 1 > gpurun_out/s9_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc_bwd8 -s 3 -c 1 -f -o gpurun_out/r1_ncu_attn_bwd8 python tools/profile_step.py --steps 1 > gpurun_out/s9_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 6 -c 1 -f -o gpurun_out/r1_ncu_qkv_gemm python tools/gemm_probe.py qkv 10 > gpurun_out/s9_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
