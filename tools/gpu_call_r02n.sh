set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02n_bench_4gpu.json 2> gpurun_out/r02n_bench_4gpu.err; echo "bench4 rc=$?"
tail -3 gpurun_out/r02n_bench_4gpu.err
