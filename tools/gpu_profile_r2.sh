cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2 flush > gpurun_out/ncu_bench_r2.log 2>&1
IDTO_SUBSTREAMS=1 IDTO_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_partials_chain|k_kkt_v3|k_partials_path|k_assemble|k_tau_chain|k_gm_matvec" -s 12 -c 7 -o gpurun_out/r2_final -f python tools/simple_steps.py 4 > gpurun_out/ncu_full_r2.log 2>&1
tail -2 gpurun_out/ncu_full_r2.log
