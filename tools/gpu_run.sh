cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/r3f_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
IDTO_NO_ZEROCOPY=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3f_bench_nozc.json 2> gpurun_out/r3f_bench_nozc.err
cat gpurun_out/r3f_bench.json gpurun_out/r3f_bench_nozc.json | python tools/benchfmt.py | tail -n 6
timeout 300 python tools/e2e_breakdown.py > gpurun_out/r3f_e2e_breakdown.log 2>&1; tail -n 8 gpurun_out/r3f_e2e_breakdown.log
