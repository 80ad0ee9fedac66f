set -x
cd $GRAFT_REPO_ROOT
# every launch of OUR kernels (and the cub scans between them) of the bench command, with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(void )?(<unnamed>::)?(fj::)?k_|DeviceScan" -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-aux > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
grep -c "k_fj" gpurun_out/r02_launches_bench_c2.csv
