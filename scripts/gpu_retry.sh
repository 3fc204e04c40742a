#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> <outfile> <command...>   — retries while the pod is busy.  An attempt is only made
# from a clean, committed tree (the snapshot gpurun takes must be a consistent state: built libraries newer than their sources).
T=$1; OUT=$2; shift 2
for i in $(seq 1 60); do
  if [ -n "$(git -C /root/repo status --porcelain --untracked-files=no)" ] || [ -e /root/repo/.editing ]; then sleep 45; continue; fi
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  if ! grep -q "status=transient" $OUT; then exit 0; fi
  sleep 45
done
