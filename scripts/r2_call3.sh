#!/bin/bash
# Round 2, GPU call 3: the whole GPU suite, then the launch-configuration sweeps (scan kernel teams; build / probe / send /
# hash-aggregate kernels), then the bench line.
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log )
tail -15 $O/r2d_pytest.log
( timeout 700 python scripts/sweep_teams.py 1e8 wide,narrow > $O/r2d_sweep_teams.jsonl 2> $O/r2d_sweep_teams.err; echo "sweep rc=$?" >> $O/r2d_sweep_teams.err )
cut -c1-200 $O/r2d_sweep_teams.jsonl; tail -3 $O/r2d_sweep_teams.err
( timeout 900 python scripts/sweep_np.py 1e8 build,probe,motion,groupby > $O/r2d_sweep_np.jsonl 2> $O/r2d_sweep_np.err; echo "sweep rc=$?" >> $O/r2d_sweep_np.err )
cut -c1-220 $O/r2d_sweep_np.jsonl; tail -3 $O/r2d_sweep_np.err
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2d_bench.json 2> $O/r2d_bench.err; echo "bench rc=$?" >> $O/r2d_bench.err )
tail -c 3000 $O/r2d_bench.json; tail -5 $O/r2d_bench.err
