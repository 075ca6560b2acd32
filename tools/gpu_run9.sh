cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2h_tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload allegro_hand --method forward > gpurun_out/r2h_bench_allegro.json 2>> gpurun_out/r2h_bench.err
IDTO_MAX_ACTIVE_PAIRS=32 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload allegro_hand --method forward > gpurun_out/r2h_bench_allegro32.json 2>> gpurun_out/r2h_bench.err
