cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/kkt_check.py > gpurun_out/r3c_kkt_check.log 2>&1; echo "kkt_check rc=$?"
grep "^solver" gpurun_out/r3c_kkt_check.log | cut -c1-200
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/r3c_tests.log
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xn/libidto_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r3c_tests_nopipe.log 2>&1; echo "tests nopipe rc=$?"
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xt/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 2>&1 | grep -E "kkt3|ms/step" | tail -n 40 > gpurun_out/r3c_kkt_timing.log
grep "kkt3 dir 0" gpurun_out/r3c_kkt_timing.log | tail -n 3
grep "kkt3 sub dir 0" gpurun_out/r3c_kkt_timing.log | tail -n 3
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xn/libidto_b200.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3c_bench_nopipe.json 2> gpurun_out/r3c_bench_nopipe.err
cat gpurun_out/r3c_bench.json gpurun_out/r3c_bench_nopipe.json | python tools/benchfmt.py | tail -n 6
