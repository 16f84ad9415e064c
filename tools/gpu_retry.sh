#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <logfile> '<command>' [gpurun extra args]: retries while the pod answers busy (exit code 3)
T=$1; LOG=$2; CMD=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then echo "done rc=$rc try=$i" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$LOG"; exit 3
