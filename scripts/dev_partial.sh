#!/bin/bash
# default (adaptive) configuration on both layouts, then the PARTIAL-stage plan (8 value slots) under several configurations
for tab in wide narrow; do
  n=100000000; [ $tab = narrow ] && n=200000000
  echo "== default $tab"; timeout 200 python scripts/dev_q1.py $n $tab 2>&1 | grep "iter 5"
done
for cfg in "" "9,4" "11,3" "13,3" "12,3" "14,2"; do
  echo "== partial wide cfg='$cfg'"
  GGB200_PLAN_CACHE=0 GGB200_PRIV_CONFIG="$cfg" timeout 200 python scripts/dev_q1.py 100000000 wide x partial 2>&1 | grep "iter 5"
done
