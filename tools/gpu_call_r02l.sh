set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_gputests.log
tail -25 gpurun_out/r02l_gputests.log
timeout 300 python tools/ensemble_bench.py 100000 > gpurun_out/r02l_ensemble.txt 2>&1; cat gpurun_out/r02l_ensemble.txt
DW_NO_TRIM=1 timeout 300 python tools/ensemble_bench.py 100000 > gpurun_out/r02l_ensemble_notrim.txt 2>&1; cat gpurun_out/r02l_ensemble_notrim.txt
timeout 100 python tools/fused_bench.py quick > gpurun_out/r02l_fused.txt 2>&1; cat gpurun_out/r02l_fused.txt
