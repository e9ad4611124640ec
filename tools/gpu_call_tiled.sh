set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tiled_parity.py -m gpu -q 2>&1 | tail -3
timeout 200 python tools/tiled_bench.py 16384 2>&1 | tail -2
DW_TILED_ONE_TILE=1 timeout 200 python tools/tiled_bench.py 16384 2>&1 | tail -2
timeout 200 python - <<'PY'
import sys; sys.path.insert(0, '.')
import os
from therldaisyworld_b200.banded import BandedDaisyWorld
# one band of 2048 rows of a 16384-wide world is not constructible stand-alone; time the stencil on square worlds instead
for N in (4096, 16384):
    for env in ("", "1"):
        if env: os.environ["DW_TILED_ONE_TILE"] = "1"
        else: os.environ.pop("DW_TILED_ONE_TILE", None)
        w = BandedDaisyWorld(N, 64)
        w.reset_on_device(seed=1); w.run(3, "greedy")
        print(N, "one-tile" if env else "persistent", f"{w.band.time_stencil(30):.1f} us per stencil launch", flush=True)
        del w
PY
