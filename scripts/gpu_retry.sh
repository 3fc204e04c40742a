#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> <outfile> <command...>   — retries while the pod is busy
T=$1; OUT=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  if ! grep -q "status=transient" $OUT; then exit 0; fi
  sleep 60
done
