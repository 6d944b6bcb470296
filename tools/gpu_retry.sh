#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'   -- retries gpurun while the pod answers busy/transient (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpu_retry.out 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpu_retry.out || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
cat /tmp/gpu_retry.out
exit $rc
