#!/bin/bash
# tools/gpurun_retry.sh <timeout> <command> — gpurun, retried while the pod answers "busy / draining" (nothing is charged for those)
to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout $to -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|rc=3\|retry in a few minutes"; then sleep 150; continue; fi
  break
done
