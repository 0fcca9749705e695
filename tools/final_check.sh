#!/bin/bash
# last check of the round on the final commit: full GPU suite, smoke, refreshed 2-GPU line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/fc_pytest.log; tail -3 gpurun_out/fc_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fc_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/fc_smoke.log; tail -2 gpurun_out/fc_smoke.log
