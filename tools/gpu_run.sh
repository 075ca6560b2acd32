cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for B in 64 8; do
for ls in sweep cyclic_reduction; do
timeout 300 python bench.py --no-cpu-baseline --batch $B --linear-solver $ls --steps 20 --warmup 3 > gpurun_out/r3o_bench_${ls}_b$B.json 2> gpurun_out/r3o_bench_${ls}_b$B.err
echo "batch $B $ls"; tail -n 1 gpurun_out/r3o_bench_${ls}_b$B.json | python tools/benchfmt.py 2>/dev/null | head -n 2; tail -n 2 gpurun_out/r3o_bench_${ls}_b$B.err
done; done
