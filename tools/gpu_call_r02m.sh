set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/segment_bench.py 100000 > gpurun_out/r02m_segments.txt 2>&1; cat gpurun_out/r02m_segments.txt
timeout 200 python tools/segment_bench.py 2000 >> gpurun_out/r02m_segments.txt 2>&1; tail -5 gpurun_out/r02m_segments.txt
