#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged)
# usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "--- transient, retry $i in 90 s"
  sleep 90
done
