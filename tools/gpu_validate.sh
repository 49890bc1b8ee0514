#!/bin/bash
# One GPU-box call that validates a build: every GPU test, the driver's smoke(), the default bench line and the CPU arm.
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'        (about 2.5 minutes of box time)
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^E   *+" | tail -15) > gpurun_out/validate_pytest.log 2>&1
tail -4 gpurun_out/validate_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1
tail -1 gpurun_out/validate_smoke.log
timeout 600 python bench.py > gpurun_out/validate_bench.log 2>&1
tail -1 gpurun_out/validate_bench.log | cut -c1-400
