#!/bin/bash
# One GPU-box call that refreshes the evidence under profiles/: the per-launch list of the C2 step, ncu --set full captures
# of the LayerNorm-backward, attention-backward and fused-QKV GEMM kernels, and the single-kernel probe tables.
#   gpurun --timeout 1800 -- 'bash tools/gpu_profile.sh'          (about 3 minutes of box time)
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py --steps 1 > gpurun_out/profile_step.log 2>&1
python tools/profile_step.py --summarise gpurun_out/launches_step.csv > gpurun_out/launches_step.md
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:add_ln_bwd_pipe -s 3 -c 1 -f -o gpurun_out/ncu_ln_bwd_pipe python tools/profile_step.py --steps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc_bwd8 -s 3 -c 1 -f -o gpurun_out/ncu_attn_bwd8 python tools/profile_step.py --steps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 6 -c 1 -f -o gpurun_out/ncu_qkv_gemm python tools/gemm_probe.py qkv 10 > /dev/null 2>&1
for what in gemm ln attn mul; do timeout 600 python tools/op_probe.py $what > gpurun_out/probe_$what.log 2>&1; done
ls -la gpurun_out
