#!/bin/bash
# rows per lane x (consumer warps, ring stages) on the Q1 scan kernel (NVRTC path)
for cfg in "2 8,5" "2 10,5" "2 12,4" "2 20,3" "1 16,4" "1 20,3" "1 18,3" "1 22,3"; do
  set -- $cfg
  for tab in wide narrow; do
    n=100000000; [ $tab = narrow ] && n=200000000
    echo "== rows_per_lane=$1 cfg=$2 $tab"
    GGB200_PLAN_CACHE=0 GGB200_ROWS_PER_LANE=$1 GGB200_PRIV_CONFIG=$2 timeout 200 python scripts/dev_q1.py $n $tab 2>&1 | grep "iter 5"
  done
done
