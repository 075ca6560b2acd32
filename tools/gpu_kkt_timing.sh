cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_timing/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | grep -E "kkt3|ms/step" | tail -14 > gpurun_out/kkt_timing.log
