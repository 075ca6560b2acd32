cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" _xB _xC; do
  IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib$v/libidto_b200.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench$v.json 2> gpurun_out/r2e_bench$v.err
done
