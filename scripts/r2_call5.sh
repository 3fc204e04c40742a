#!/bin/bash
# Round 2, GPU call 5 (first call of the third session): the whole GPU suite (no stop at the first failure), the full bench
# line with every secondary, then the ncu evidence of the bench command: launch list + one `--set full` capture of the
# scan kernel.
mkdir -p gpurun_out
O=gpurun_out
T=r2e
( timeout 1200 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log )
tail -25 $O/${T}_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?" >> $O/${T}_bench.err )
tail -c 1500 $O/${T}_bench.json; tail -5 $O/${T}_bench.err
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --secondary none"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $BENCH > $O/${T}_ncu_launches.log 2>&1
grep -c gg_ $O/${T}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gg_jit_scanagg --launch-skip 3 -c 1 -o $O/${T}_prof_scanagg $BENCH > $O/${T}_ncu_scanagg.log 2>&1
ls -la $O | grep ${T}_
