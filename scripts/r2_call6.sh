#!/bin/bash
# Round 2, GPU call 6: the test files call 5 never reached (it hung in the first test whose teams wrap the ring: a team waited for
# a phase that had already passed), the staged AOCS scan, the bench line, the random-access rates.  Every step under its own
# timeout; a hang stops the script (the steps after it would only queue behind a wedged context).
mkdir -p gpurun_out
O=gpurun_out
T=r2f
timeout -s KILL 420 python -m pytest tests/test_gpu_scanagg.py tests/test_gpu_zz_reference_goldens.py tests/test_gpu_sort.py -q -x > $O/${T}_pytest_tail.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${T}_pytest_tail.log; tail -15 $O/${T}_pytest_tail.log
if [ $rc -ge 124 ]; then echo "hang: stopping"; exit 1; fi
timeout -s KILL 60 build/gather_peak > $O/${T}_gather_peak.json 2> $O/${T}_gather_peak.err; cat $O/${T}_gather_peak.json
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
echo "bench rc=$?" >> $O/${T}_bench.err
tail -c 1200 $O/${T}_bench.json; tail -5 $O/${T}_bench.err
