set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "agg" 2>&1 | tail -4
AGGBENCH=q1,q1_s3,q1_twobuf timeout 600 python tools/aggbench.py 2>&1 | tail -6
