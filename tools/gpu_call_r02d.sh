set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_gputests.log
tail -30 gpurun_out/r02d_gputests.log
echo "=== fused MLP" > gpurun_out/r02d_mlp.txt
timeout 200 python tools/fused_bench.py mlp >> gpurun_out/r02d_mlp.txt 2>&1
echo "=== per-step MLP (DW_MLP_UNFUSED=1)" >> gpurun_out/r02d_mlp.txt
DW_MLP_UNFUSED=1 timeout 200 python tools/fused_bench.py mlp >> gpurun_out/r02d_mlp.txt 2>&1
cat gpurun_out/r02d_mlp.txt
timeout 100 python - > gpurun_out/r02d_f32.txt 2>&1 <<'PY'
import sys, json
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.materialise_bench(0, bench.load_peaks()), indent=1))
PY
cat gpurun_out/r02d_f32.txt
