#!/bin/bash
# Round 2, GPU call 2: full GPU test suite + the new bench.py (executor-node surface, parity gate, secondary block) at N=1
# + launch-configuration sweep of the team-scheduled scan kernel.
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log )
tail -5 $O/r2b_pytest.log
( timeout 900 python scripts/sweep_teams.py 1e8 wide,narrow > $O/r2c_sweep.jsonl 2> $O/r2c_sweep.err; echo "sweep rc=$?" >> $O/r2c_sweep.err )
cat $O/r2c_sweep.jsonl | cut -c1-220; tail -3 $O/r2c_sweep.err
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?" >> $O/r2b_bench.err )
tail -c 7000 $O/r2b_bench.json; tail -20 $O/r2b_bench.err
