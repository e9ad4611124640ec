#!/bin/bash
# parity of the default build, then the fused micro-bench for every built variant
mkdir -p gpurun_out
python -W ignore -m pytest tests/test_gpu_run_parity.py tests/test_gpu_fused_internals.py -x -q 2>&1 | tail -5
python tools/fused_bench.py quick
for lib in therldaisyworld_b200/libdaisyworld_b200.*.so; do
  DW_LIB=$lib python tools/fused_bench.py quick
done
