cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2p_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench_$tag.json 2> gpurun_out/r2p_bench_$tag.err; }
run graph A=1
run nograph IDTO_GRAPH=0
run graph2 A=1
run nograph2 IDTO_GRAPH=0
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2p_e2e_breakdown.log 2>&1
