cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mpc_shell.py tests/test_more_options_gpu.py tests/test_headline_parity.py -x -q -m gpu > gpurun_out/r3k_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/r3k_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; echo "bench rc=$?"
cat gpurun_out/r3k_bench.json | python tools/benchfmt.py | tail -n 2
timeout 300 python tools/e2e_breakdown.py > gpurun_out/r3k_e2e_breakdown.log 2>&1; tail -n 8 gpurun_out/r3k_e2e_breakdown.log
