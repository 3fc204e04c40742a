#!/bin/bash
# Round 2, GPU call 7: A/B of the headline scan kernel (default build, without the snapshot rule, teams on the slots' barriers),
# the join tests over the new probe loop (two probe steps per round trip, distinct-key shortcut), the probe's launch
# configurations.  Every step under its own timeout.
mkdir -p gpurun_out
O=gpurun_out
T=r2g
: > $O/${T}_ab_scan.jsonl
for lib in "" build/ab_nomvcc/libggb200.so build/ab_slotbar/libggb200.so; do
  GGB200_DEVLIB=$lib timeout -s KILL 150 python scripts/ab_scan.py 1e8 >> $O/${T}_ab_scan.jsonl 2>> $O/${T}_ab_scan.err
done
cat $O/${T}_ab_scan.jsonl
timeout -s KILL 400 python -m pytest tests/test_gpu_join.py tests/test_gpu_random_joins.py -q -x > $O/${T}_pytest_join.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${T}_pytest_join.log; tail -6 $O/${T}_pytest_join.log
if [ $rc -ge 124 ]; then echo "hang: stopping"; exit 1; fi
SWEEP_PROBE="-;-/u0;15,3,0,2;14,3,0,2;24,4,0;28,3,0;20,4,0" timeout -s KILL 300 python scripts/sweep_np.py 1e8 probe > $O/${T}_sweep_probe.jsonl 2> $O/${T}_sweep_probe.err
echo "sweep rc=$?"; cut -c1-250 $O/${T}_sweep_probe.jsonl; tail -3 $O/${T}_sweep_probe.err
