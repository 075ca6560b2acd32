cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in timing xserial xsmem xsmemserial; do
  IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_$v/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | grep -E "kkt3|ms/step|wall" | tail -9 > gpurun_out/r2d_kkt_$v.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_headline_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2d_tests.log
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xsmem/libidto_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_headline_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2d_tests_smem.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
