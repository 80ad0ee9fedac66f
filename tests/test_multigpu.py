"""-m gpu test of the N > 1 path on real GPUs: launches tests/multigpu_worker.py with one process per GPU (2 when the
box has at least two, else the same program with a single rank so that the code path still runs)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_push_shuffle_join_agg_q3_across_ranks():
    import torch
    n = min(torch.cuda.device_count(), int(os.environ.get("GSQL_TEST_GPUS", "2")))
    assert n >= 1, "no CUDA device"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 2000), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    if r.returncode != 0:  # keep the whole transcript where a gpurun call brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "multigpu_worker.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-6000:])
