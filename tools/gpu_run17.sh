cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2o_e2e_breakdown.log 2>&1
