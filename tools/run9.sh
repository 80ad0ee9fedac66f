set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_pipelines_gpu.py tests/test_jni_boundary.py -m gpu -x -q -k "agg or q3 or two_phase or jni or Jni" 2>&1 | tail -4
AGGBENCH=c5,c5_nopart timeout 600 python tools/aggbench.py 2>&1 | tail -6
