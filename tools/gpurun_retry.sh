#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   — retries while the pod answers "transient" (no slot; nothing charged)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
