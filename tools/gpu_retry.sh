#!/bin/bash
# tools/gpu_retry.sh <logfile> <timeout> <command...>: retries a gpurun call while the pod answers "transient" / busy
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" $log && ! grep -q "status=ok" $log; then sleep 120; else break; fi
done
