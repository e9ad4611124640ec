set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_gputests.log
tail -5 gpurun_out/r02h_gputests.log
timeout 200 python tools/fused_bench.py mlp > gpurun_out/r02h_mlp.txt 2>&1; cat gpurun_out/r02h_mlp.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02h_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02h_bench_2gpu_ref.json 2> gpurun_out/r02h_bench_2gpu_ref.err; echo "ref2 rc=$?"
