set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_1gpu.json 2> gpurun_out/r02k_bench_1gpu.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02k_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -c 1 -o gpurun_out/r02k_fused_final -f python tools/profile_target.py 1000 385 64 > gpurun_out/ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forward_f32 -s 1 -c 1 -o gpurun_out/r02k_f32 -f python tools/f32_profile_target.py 4000 > gpurun_out/ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_n64_persist -c 1 -o gpurun_out/r02k_mlp -f python tools/mlp_profile_target.py 1000 129 > gpurun_out/ncu_c.log 2>&1; echo "ncu c rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile4 -c 1 -o gpurun_out/r02k_tile4_n96 -f python tools/profile_target.py 888 129 96 > gpurun_out/ncu_d.log 2>&1; echo "ncu d rc=$?"
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log
