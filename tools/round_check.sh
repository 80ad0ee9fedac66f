#!/bin/bash
# One GPU-box call: full GPU test suite, the 1-GPU bench line, the ncu launch list of the same command, one --set full capture.
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:"k_|gsql|Device" --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fj_probe|k_fj_scatter|k_fj_hist|k_fj_build_part" -s 5 -c 5 -o gpurun_out/prof_r01k python tools/steptime.py 1.0 > gpurun_out/ncu_k.log 2>&1; echo "ncu full rc=$?"
