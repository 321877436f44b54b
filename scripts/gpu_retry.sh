#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> <command...>  -- retries while the pod answers "transient" (nothing charged)
t=$1; shift
for i in $(seq 1 25); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out" | tail -40; exit 0
done
echo "gave up after 25 transient answers"
