cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mpc_shell.py tests/test_headline_parity.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2k_tests.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 600 python bench.py --steps 50 --warmup 5 --impl reference > gpurun_out/r2k_bench_ref.json 2>> gpurun_out/r2k_bench.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --method forward > gpurun_out/r2k_bench_fwd.json 2>> gpurun_out/r2k_bench.err
