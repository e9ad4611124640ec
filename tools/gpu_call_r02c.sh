set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_gputests.log
tail -30 gpurun_out/r02c_gputests.log
for v in "" aw7 aw3; do
  if [ -z "$v" ]; then unset DW_LIB; else export DW_LIB=$GRAFT_REPO_ROOT/therldaisyworld_b200/libdaisyworld_b200.$v.so; fi
  echo "=== variant ${v:-base}" >> gpurun_out/r02c_variants.txt
  timeout 120 python tools/fused_bench.py quick >> gpurun_out/r02c_variants.txt 2>&1
done
unset DW_LIB
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "bench rc=$?"
