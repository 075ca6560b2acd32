cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_direct.out 2> gpurun_out/r2q_direct.err
echo "exit $?" >> gpurun_out/r2q_direct.err; ls -la gpurun_out/r2q_direct.out >> gpurun_out/r2q_direct.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/r2q_direct_strong.out 2>> gpurun_out/r2q_direct.err
ls -la gpurun_out/r2q_direct_strong.out >> gpurun_out/r2q_direct.err
