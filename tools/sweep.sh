#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_operators_gpu.py -x -q -k "agg or Agg" 2>&1 | tail -5
timeout 300 python tools/aggbench.py 1.0 2>&1 | tail -2
