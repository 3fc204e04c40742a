#!/bin/bash
# Round 2, GPU call 4: the whole GPU suite (every test, no stop at the first failure), the launch-configuration sweeps of the
# build / probe / send / hash-aggregate kernels (one process per operator: a faulting configuration leaves a sticky CUDA error
# behind), the bench line, and the ncu evidence of the bench command: launch list + one `--set full` capture of the scan kernel.
mkdir -p gpurun_out
O=gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log )
tail -12 $O/r2e_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?" >> $O/r2e_bench.err )
tail -c 2500 $O/r2e_bench.json; tail -5 $O/r2e_bench.err
: > $O/r2e_sweep_np.jsonl
for op in build probe motion groupby; do
    ( timeout 420 python scripts/sweep_np.py 1e8 $op >> $O/r2e_sweep_np.jsonl 2> $O/r2e_sweep_np_$op.err; echo "sweep $op rc=$?" >> $O/r2e_sweep_np_$op.err )
    tail -2 $O/r2e_sweep_np_$op.err
done
cut -c1-200 $O/r2e_sweep_np.jsonl
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --secondary none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2e_launches.csv $BENCH > $O/r2e_ncu_launches.log 2>&1
grep -c gg_ $O/r2e_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gg_jit_scanagg --launch-skip 3 -c 1 -o $O/r2e_prof_scanagg $BENCH > $O/r2e_ncu_scanagg.log 2>&1
ls -la $O | grep r2e_
