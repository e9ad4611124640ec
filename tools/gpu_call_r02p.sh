set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_diag_feed.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python - > gpurun_out/r02p_f32.txt 2>&1 <<'PY'
import sys, json
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.materialise_bench(0, bench.load_peaks()), indent=1))
PY
cat gpurun_out/r02p_f32.txt | head -30
