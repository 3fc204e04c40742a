#!/bin/bash
# register-resident accumulators: default configuration, a few (warps, stages) alternatives, and the old layout
for cfg in "" "20,4" "18,4" "20,3"; do
  echo "== wide normal cfg='$cfg'"; GGB200_PRIV_CONFIG="$cfg" timeout 200 python scripts/dev_q1.py 100000000 wide 2>&1 | grep "iter 5"
done
echo "== wide normal, no register slots"; GGB200_REG_SLOTS=0 GGB200_PLAN_CACHE=0 timeout 200 python scripts/dev_q1.py 100000000 wide 2>&1 | grep "iter 5"
for cfg in "" "16,4" "20,3"; do
  echo "== wide partial cfg='$cfg'"; GGB200_PRIV_CONFIG="$cfg" timeout 200 python scripts/dev_q1.py 100000000 wide x partial 2>&1 | grep "iter 5"
done
echo "== narrow normal"; timeout 200 python scripts/dev_q1.py 200000000 narrow 2>&1 | grep "iter 5"
echo "== narrow normal 20,4"; GGB200_PRIV_CONFIG="20,4" timeout 200 python scripts/dev_q1.py 200000000 narrow 2>&1 | grep "iter 5"
