#!/bin/bash
# Round 2, GPU call 2: full GPU test suite + the new bench.py (executor-node surface, parity gate, secondary block) at N=1.
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log )
tail -5 $O/r2b_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?" >> $O/r2b_bench.err )
tail -c 6000 $O/r2b_bench.json; tail -20 $O/r2b_bench.err
