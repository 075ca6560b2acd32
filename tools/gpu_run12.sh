cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2j_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_$tag.json 2> gpurun_out/r2j_bench_$tag.err; }
run base A=1
run matvec IDTO_TRUST_MATVEC=1
