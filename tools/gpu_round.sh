#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the bench command, full captures of the top kernels.
set -x
TAG=${1:-r1d}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -W ignore -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -c 1 -f -o gpurun_out/prof_fused_$TAG \
    python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tiled_step -s 4 -c 1 -f -o gpurun_out/prof_tiled_$TAG \
    python tools/tiled_bench.py 16384 > gpurun_out/ncu_full_tiled_$TAG.log 2>&1; echo "ncu tiled rc=$?"
ls -la gpurun_out | tail -15
