set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_gputests.log
tail -25 gpurun_out/r02i_gputests.log
timeout 300 python tools/n_sweep.py > gpurun_out/r02i_n_sweep.txt 2>&1; cat gpurun_out/r02i_n_sweep.txt
timeout 200 python tools/fuzz_fast_path.py 80 777 13,18,33,61,75,150,20,96 2>&1 | tail -3 > gpurun_out/r02i_fuzz.txt; cat gpurun_out/r02i_fuzz.txt
timeout 100 python tools/fused_bench.py quick > gpurun_out/r02i_fused.txt 2>&1; cat gpurun_out/r02i_fused.txt
