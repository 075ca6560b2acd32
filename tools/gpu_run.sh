cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/r3g_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"
IDTO_PATH_CONCURRENT=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3g_bench_seq.json 2> gpurun_out/r3g_bench_seq.err
cat gpurun_out/r3g_bench.json gpurun_out/r3g_bench_seq.json | python tools/benchfmt.py | tail -n 6
