#!/bin/bash
# usage: tools/gpurun_retry.sh <log> [gpurun args...] -- '<command>'
# retries while the pod has no free slot (exit code 3 / status=transient: nothing charged)
log=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
