set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_fused_internals.py tests/test_gpu_run_parity.py -m gpu -x -q > gpurun_out/r02j_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_gputests.log
tail -5 gpurun_out/r02j_gputests.log
timeout 300 python tools/n_sweep.py > gpurun_out/r02j_n_sweep.txt 2>&1; cat gpurun_out/r02j_n_sweep.txt
