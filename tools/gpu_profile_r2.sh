# Round-2 evidence run (1 GPU): bench lines of the build, then the two ncu passes of B200_PROFILING.md.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/r3_bench_central.json 2> gpurun_out/r3_bench_central.err
python bench.py --l2 flush --no-cpu-baseline > gpurun_out/r3_bench_central_flush.json 2>/dev/null
python bench.py --method forward --no-cpu-baseline > gpurun_out/r3_bench_forward.json 2>/dev/null
python bench.py --impl reference > gpurun_out/r3_bench_reference.json 2>/dev/null
python bench.py --workload hopper --no-cpu-baseline > gpurun_out/r3_bench_hopper.json 2>/dev/null
python bench.py --workload allegro_hand --no-cpu-baseline > gpurun_out/r3_bench_allegro.json 2>/dev/null
python bench.py --scaling strong --no-cpu-baseline > gpurun_out/r3_bench_strong1.json 2>/dev/null
IDTO_PATH_CONCURRENT=0 python bench.py --no-cpu-baseline > gpurun_out/r3_bench_central_seq.json 2>/dev/null
IDTO_NO_ZEROCOPY=1 python bench.py --no-cpu-baseline > gpurun_out/r3_bench_central_nozc.json 2>/dev/null
python tools/e2e_breakdown.py > gpurun_out/r3_e2e_breakdown.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2 flush > gpurun_out/ncu_bench_r2.log 2>&1
IDTO_SUBSTREAMS=1 IDTO_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_partials_chain|k_kkt_v3|k_partials_path|k_assemble|k_tau_chain|k_gm_matvec" -s 12 -c 7 -o gpurun_out/r2_final -f python tools/simple_steps.py 4 > gpurun_out/ncu_full_r2.log 2>&1
tail -n 2 gpurun_out/ncu_full_r2.log
for f in gpurun_out/r3_bench_*.json; do echo $f; tail -n 1 $f | python tools/benchfmt.py | head -n 1; done
