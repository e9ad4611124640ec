set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gputests.log
tail -5 gpurun_out/r02e_gputests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python tools/n_sweep.py > gpurun_out/r02e_n_sweep.txt 2>&1
cat gpurun_out/r02e_n_sweep.txt
timeout 200 python tools/es_bench.py > gpurun_out/r02e_es.txt 2>&1; cat gpurun_out/r02e_es.txt
