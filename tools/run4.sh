set -x
cd $GRAFT_REPO_ROOT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_agg_consume|k_xchg_push|k_push_hist|k_scan|k_agg_part_scatter|k_agg_reg" -c 12 -f -o gpurun_out/r02_kernels python tools/kprof.py > gpurun_out/kprof.log 2>&1
tail -5 gpurun_out/kprof.log
ls -la gpurun_out/*.ncu-rep
