set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -5
grep -v "^\[W\|^W0" gpurun_out/multigpu_worker.log | grep -v "^  File\|^    " | tail -30
