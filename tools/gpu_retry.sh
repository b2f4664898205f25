#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'  — retries gpurun while the pod answers busy (nothing is charged then)
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpu_retry.out 2>&1
  if ! grep -q "status=transient" /tmp/gpu_retry.out; then break; fi
  sleep 45
done
tail -n 80 /tmp/gpu_retry.out
