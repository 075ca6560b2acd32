# Round-2 evidence run (1 GPU): tests, smoke, bench lines of the build, then the two ncu passes of B200_PROFILING.md.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r4_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r4_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r4_smoke.log
IDTO_GRAPH=0 IDTO_PATH_CONCURRENT=0 IDTO_NO_ZEROCOPY=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_mpc_shell.py tests/test_more_options_gpu.py -x -q -m gpu > gpurun_out/r4_tests_switches_off.log 2>&1; echo "tests (graph, side-by-side, zero-copy off) rc=$?"; tail -n 1 gpurun_out/r4_tests_switches_off.log
python bench.py > gpurun_out/r4_bench_central.json 2> gpurun_out/r4_bench_central.err
python bench.py --impl reference > gpurun_out/r4_bench_reference.json 2>/dev/null
python bench.py --l2 flush --no-cpu-baseline > gpurun_out/r4_bench_central_flush.json 2>/dev/null
python bench.py --method forward --no-cpu-baseline > gpurun_out/r4_bench_forward.json 2>/dev/null
python bench.py --workload hopper --no-cpu-baseline > gpurun_out/r4_bench_hopper.json 2>/dev/null
python bench.py --workload hopper --method forward --no-cpu-baseline > gpurun_out/r4_bench_hopper_forward.json 2>/dev/null
python bench.py --workload allegro_hand --no-cpu-baseline > gpurun_out/r4_bench_allegro.json 2>/dev/null
python bench.py --workload allegro_hand --method forward --no-cpu-baseline > gpurun_out/r4_bench_allegro_forward.json 2>/dev/null
python bench.py --scaling strong --no-cpu-baseline > gpurun_out/r4_bench_strong1.json 2>/dev/null
python bench.py --linear-solver cyclic_reduction --no-cpu-baseline > gpurun_out/r4_bench_cyclic_reduction.json 2>/dev/null
IDTO_PATH_CONCURRENT=0 python bench.py --no-cpu-baseline > gpurun_out/r4_bench_central_seq.json 2>/dev/null
IDTO_NO_ZEROCOPY=1 python bench.py --no-cpu-baseline > gpurun_out/r4_bench_central_nozc.json 2>/dev/null
IDTO_GRAPH=0 python bench.py --no-cpu-baseline > gpurun_out/r4_bench_central_nograph.json 2>/dev/null
python tools/e2e_breakdown.py > gpurun_out/r4_e2e_breakdown.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2 flush > gpurun_out/ncu_bench_r2.log 2>&1
IDTO_SUBSTREAMS=1 IDTO_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_partials_chain|k_kkt_v3|k_partials_path|k_assemble|k_tau_chain|k_gm_matvec" -s 12 -c 7 -o gpurun_out/r2_final -f python tools/simple_steps.py 4 > gpurun_out/ncu_full_r2.log 2>&1
tail -n 2 gpurun_out/ncu_full_r2.log
