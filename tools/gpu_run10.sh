cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
IDTO_SUBSTREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_partials_chain|k_kkt_v3|k_partials_path|k_assemble" -s 8 -c 4 -o gpurun_out/r2_full -f python tools/simple_steps.py 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
