set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_gputests.log
tail -40 gpurun_out/r02f_gputests.log
timeout 200 python tools/fused_bench.py mlp > gpurun_out/r02f_mlp.txt 2>&1; cat gpurun_out/r02f_mlp.txt
timeout 200 python tools/es_bench.py > gpurun_out/r02f_es.txt 2>&1; cat gpurun_out/r02f_es.txt
