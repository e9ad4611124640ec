python -W ignore -m pytest tests/test_gpu_run_parity.py tests/test_gpu_fused_internals.py -x -q 2>&1 | tail -3
for kc in 8 16 32 64; do echo KC=$kc; DW_PERSIST_KC=$kc python tools/fused_bench.py quick; done
