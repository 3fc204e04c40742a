#!/bin/bash
for cfg in "20,5" "18,5" "16,5" "20,4"; do
  echo "== wide normal cfg='$cfg'"; GGB200_PRIV_CONFIG="$cfg" timeout 200 python scripts/dev_q1.py 100000000 wide 2>&1 | grep "iter [45]"
done
echo "== narrow partial (default rule)"; timeout 200 python scripts/dev_q1.py 200000000 narrow x partial 2>&1 | grep "iter [45]"
