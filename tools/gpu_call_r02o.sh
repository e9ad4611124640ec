set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_gputests.log
tail -6 gpurun_out/r02o_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02o_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02o_bench_1gpu.json 2> gpurun_out/r02o_bench_1gpu.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -s 1 -c 1 -o gpurun_out/r02o_mlp -f python tools/mlp_profile_target.py 1000 129 > gpurun_out/ncu_c.log 2>&1; echo "ncu mlp rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forward_f32 -s 1 -c 1 -o gpurun_out/r02o_f32 -f python tools/f32_profile_target.py 4000 > gpurun_out/ncu_b.log 2>&1; echo "ncu f32 rc=$?"
