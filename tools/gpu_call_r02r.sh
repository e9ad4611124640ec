set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tiled_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02r_bench_2gpu.json 2> gpurun_out/r02r_bench_2gpu.err; echo "bench2 rc=$?"
DW_BAND_SPLIT_AGENT_KERNELS=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02r_bench_2gpu_split.json 2> gpurun_out/r02r_bench_2gpu_split.err; echo "bench2 split rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02r_bench_2gpu.json", "gpurun_out/r02r_bench_2gpu_split.json"):
    g = json.load(open(f))["giant_grid"]
    print(f, {k: g.get(k) for k in ("value", "us_per_step", "us_stencil", "us_exchange", "peer_barrier_timed_out", "error")}, g.get("state_checksum", {}).get("covers"))
PY
