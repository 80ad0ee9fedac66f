set -x
cd $GRAFT_REPO_ROOT
KPROF=agg_c5,scan,agg_q1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_agg_reg_pipe|k_agg_consume|k_scan_fast" -c 4 -f -o gpurun_out/r02_final_kernels python tools/kprof.py > gpurun_out/kprof_e.log 2>&1
tail -3 gpurun_out/kprof_e.log
ls -la gpurun_out/r02_final_kernels.ncu-rep
