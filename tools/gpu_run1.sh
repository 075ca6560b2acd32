set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_headline_parity.py 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests_old.log
timeout 1500 python -m pytest tests/test_headline_parity.py -m gpu -q 2>&1 | tail -120 > gpurun_out/r2a_gpu_tests_new.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
IDTO_B200_LIB=$GRAFT_REPO_ROOT/idto_b200/lib_timing/libidto_b200.so timeout 300 python tools/profile_step.py 2 central 64 > gpurun_out/r2a_kkt_timing.log 2>&1
tail -5 gpurun_out/r2a_gpu_tests_old.log gpurun_out/r2a_gpu_tests_new.log; cat gpurun_out/r2a_bench.json | cut -c1-600
