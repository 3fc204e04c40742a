#!/bin/bash
# Round 2, GPU call 8: the headline scan kernel without the snapshot rule compiled in (default build now), the snapshot tests
# (the pipelines compile their snapshot-rule kernel on demand), the join tests and the probe over even-aligned slot pairs.
mkdir -p gpurun_out
O=gpurun_out
T=r2h
timeout -s KILL 150 python scripts/ab_scan.py 1e8 > $O/${T}_ab_scan.jsonl 2> $O/${T}_ab_scan.err; cat $O/${T}_ab_scan.jsonl
timeout -s KILL 500 python -m pytest tests/test_gpu_join.py tests/test_gpu_random_joins.py tests/test_gpu_scanagg.py tests/test_gpu_executor.py -q -x > $O/${T}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${T}_pytest.log; tail -8 $O/${T}_pytest.log
if [ $rc -ge 124 ]; then echo "hang: stopping"; exit 1; fi
SWEEP_PROBE="-;-/u0" timeout -s KILL 200 python scripts/sweep_np.py 1e8 probe > $O/${T}_sweep_probe.jsonl 2> $O/${T}_sweep_probe.err
echo "sweep rc=$?"; cut -c1-250 $O/${T}_sweep_probe.jsonl; tail -3 $O/${T}_sweep_probe.err
