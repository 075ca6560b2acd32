cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
IDTO_SUBSTREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_r2a.csv python tools/simple_steps.py 6 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
