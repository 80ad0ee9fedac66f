#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "xchg or exchange or all_to_all or partition" 2>&1 | tail -3
for ov in 1 0; do
echo "== overlap=$ov"
GSQL_BENCH_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value']); print({k:round(v,2) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})"
done
