cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_$tag.json 2> gpurun_out/r2i_bench_$tag.err; }
run base A=1
run rt IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_xrt/libidto_b200.so
run slots4 IDTO_CHAIN_SLOTS=4
run slots6 IDTO_CHAIN_SLOTS=6
run sub1 IDTO_SUBSTREAMS=1
run sub4 IDTO_SUBSTREAMS=4
