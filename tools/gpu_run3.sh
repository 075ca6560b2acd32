cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_headline_parity.py tests/test_allegro.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c_gpu_tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_timing/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | grep -E "kkt3|ms/step" | tail -14 > gpurun_out/r2c_kkt_timing.log
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_dmma/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | grep -E "kkt3|ms/step" | tail -14 > gpurun_out/r2c_kkt_timing_dmma.log
tail -n 5 gpurun_out/r2c_gpu_tests.log
