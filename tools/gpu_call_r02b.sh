set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/micro/thr > gpurun_out/r02_micro_thr.txt 2>&1
for v in "" nosync1 nosync2 nosync12 bias3f bias0f bias33 bias3c bias30 alt altbias altbias0f altbias30; do
  if [ -z "$v" ]; then unset DW_LIB; else export DW_LIB=$GRAFT_REPO_ROOT/therldaisyworld_b200/libdaisyworld_b200.$v.so; fi
  echo "=== variant ${v:-base}" >> gpurun_out/r02_variants.txt
  timeout 120 python tools/fused_bench.py quick >> gpurun_out/r02_variants.txt 2>&1
done
for v in altbias bias3f alt; do
  export DW_LIB=$GRAFT_REPO_ROOT/therldaisyworld_b200/libdaisyworld_b200.$v.so
  echo "=== parity $v" >> gpurun_out/r02_variants_parity.txt
  timeout 300 python -m pytest tests/test_gpu_run_parity.py tests/test_gpu_fused_internals.py tests/test_gpu_tiled_parity.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_variants_parity.txt
  timeout 200 python tools/fuzz_fast_path.py 60 4242 8,16,64,96 2>&1 | tail -2 >> gpurun_out/r02_variants_parity.txt
done
unset DW_LIB
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -c 1 -o gpurun_out/r02_fused_base -f python tools/profile_target.py 1000 385 64 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
export DW_LIB=$GRAFT_REPO_ROOT/therldaisyworld_b200/libdaisyworld_b200.altbias.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -c 1 -o gpurun_out/r02_fused_altbias -f python tools/profile_target.py 1000 385 64 > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
