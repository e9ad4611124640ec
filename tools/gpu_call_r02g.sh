set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/es_profile_target.py 64 16 > gpurun_out/r02g_es.txt 2>&1
DW_POP_UNFUSED=1 python tools/es_profile_target.py 64 16 >> gpurun_out/r02g_es.txt 2>&1
cat gpurun_out/r02g_es.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02g_es_launches.csv python tools/es_profile_target.py 64 16 > gpurun_out/ncu_es.log 2>&1
python tools/launch_summary.py gpurun_out/r02g_es_launches.csv | tee gpurun_out/r02g_es_launch_summary.txt
