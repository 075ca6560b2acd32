cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_more_options_gpu.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2m_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_$tag.json 2> gpurun_out/r2m_bench_$tag.err; }
run base A=1
run p3 IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xp3/libidto_b200.so
run p4 IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xp4/libidto_b200.so
