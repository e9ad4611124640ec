set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_gputests.log
tail -5 gpurun_out/r02s_gputests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_1gpu.json 2> gpurun_out/r02s_bench_1gpu.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02s_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
