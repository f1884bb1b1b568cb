#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout_s> <command...>   -- retries gpurun while the pod has no free slot (exit 3)
log=$1; shift; to=$1; shift
for try in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  echo "[retry] try $try rc=$rc" >> "$log"
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
