#!/bin/bash
# Development aid: run a gpurun command, retrying while the pod has no free GPU slot (exit code 3 = nothing charged).
# usage: tools/gpu_retry.sh <timeout-seconds> <logfile> <command...>
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
