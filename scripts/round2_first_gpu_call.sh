#!/bin/bash
# The first GPU call of the next round, in one go (run ON the GPU box, e.g.
#   scripts/gpu_retry.sh 1500 gpurun_out/first.log 'bash scripts/round2_first_gpu_call.sh' ):
#  1. the tests that were added after round 1's last GPU minute (reference goldens, AOCS decode; -rA shows XPASS / XFAIL)
#  2. the whole GPU suite
#  3. the AOCS measurement and the other operators
#  4. ncu launch lists + one full capture each for the kernels that have none yet
# Everything lands under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_reference_goldens.py -q -rA -m gpu > gpurun_out/r2_zz.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_suite.log 2>&1
for op in aocs join groupby sort motion; do
    timeout 600 python scripts/bench_ops.py $op > gpurun_out/r2_ops_$op.json 2> gpurun_out/r2_ops_$op.err
done
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
for op in join groupby sort motion aocs; do
    timeout 900 ncu --set full --clock-control none --import-source on -c 12 -o gpurun_out/r2_prof_$op \
        python scripts/bench_ops.py $op --rows 2e7 --orders 5e6 --steps 1 --warmup 1 > gpurun_out/r2_ncu_$op.log 2>&1
done
tail -n 3 gpurun_out/r2_zz.log gpurun_out/r2_gpu_suite.log
