#!/bin/bash
# One GPU-box call: full GPU test suite, smoke, the 1-GPU bench line, the ncu launch list of the same command.
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:"k_|gsql|Device" --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/aggbench.py 1.0 2>&1 | tail -6
