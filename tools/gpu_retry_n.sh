#!/bin/bash
log=$1; shift; to=$1; shift; g=$1; shift
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus $g --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  echo "[retry] try $try rc=$rc" >> "$log"
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
