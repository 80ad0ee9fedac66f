#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then
    echo "try $i: transient, retrying" >> "$LOG.tries"
    sleep 90
    continue
  fi
  echo "try $i: rc=$rc" >> "$LOG.tries"
  exit $rc
done
exit 3
