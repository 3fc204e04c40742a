#!/bin/bash
# sweep the private-accumulator kernel configuration (consumer warps, ring stages) on both lineitem layouts
for cfg in "" "16,4" "12,5" "14,4" "18,4" "20,3" "10,5" "8,6"; do
  for tab in wide narrow; do
    n=30000000; [ $tab = narrow ] && n=60000000
    echo "== cfg='$cfg' $tab"
    GGB200_PRIV_CONFIG="$cfg" timeout 120 python scripts/dev_q1.py $n $tab 2>&1 | grep "iter [35]"
  done
done
