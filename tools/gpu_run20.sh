cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2r_bench_rotate.json 2> gpurun_out/r2r_bench_rotate.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --l2 flush > gpurun_out/r2r_bench_flush.json 2> gpurun_out/r2r_bench_flush.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload hopper --method forward > gpurun_out/r2r_bench_hopper.json 2> gpurun_out/r2r_bench_hopper.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload allegro_hand --method forward > gpurun_out/r2r_bench_allegro.json 2> gpurun_out/r2r_bench_allegro.err
