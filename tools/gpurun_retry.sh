#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy / transient (nothing charged)
to=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout $to -- "$@" 2>&1); rc=$?
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|nothing was charged\|no box or slot"; then sleep 120; continue; fi
  if [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
