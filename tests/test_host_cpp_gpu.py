"""Runs the C++ host-side operator layer (host/gsql_operators.hpp) on the GPU box: reference KATs through the mirrored
Executor / ConsumerExecutor interface in compiled code, linked only against the C-ABI."""
import subprocess

import pytest


@pytest.mark.gpu
def test_cpp_operator_layer_kats():
    import __graft_entry__ as g
    exe = g.build_host_tests()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout
    for name in ("testInnerJoin_Simple", "testLeftOuterJoin_Simple", "testHashAggSimpleCount"):
        assert f"PASS: {name}" in r.stdout


def test_cpp_operator_layer_builds_and_fails_loudly_without_gpu():
    import torch
    import __graft_entry__ as g
    exe = g.build_host_tests()
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "no usable CUDA device" in (r.stdout + r.stderr)
