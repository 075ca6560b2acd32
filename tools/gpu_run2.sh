set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_gpu_tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
IDTO_KKT_GEN=2 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_gen2.json 2>> gpurun_out/r2b_bench.err
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_timing/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | tail -12 > gpurun_out/r2b_kkt_timing.log
tail -n 8 gpurun_out/r2b_gpu_tests.log
