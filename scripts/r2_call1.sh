#!/bin/bash
# Round 2, first GPU call: operator numbers on this box + one `ncu --set full` capture of every kernel that had none in
# round 1 (join build / probe, general hash aggregate, Motion send, radix scatter, narrow scan).  Output: gpurun_out/r2a_*.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/r2a_gpu.txt 2>&1
nproc >> $O/r2a_gpu.txt
for op in join groupby sort motion aocs; do
    timeout 600 python scripts/bench_ops.py $op > $O/r2a_ops_$op.json 2> $O/r2a_ops_$op.err
done
timeout 600 python scripts/dev_q1.py 2e8 narrow > $O/r2a_q1_narrow.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:gg_ "
timeout 900 $NCU -c 6 -o $O/r2a_prof_join python scripts/bench_ops.py join --rows 4e7 --orders 1e7 --steps 1 --warmup 0 > $O/r2a_ncu_join.log 2>&1
timeout 900 $NCU -c 4 -o $O/r2a_prof_groupby python scripts/bench_ops.py groupby --rows 4e7 --steps 1 --warmup 0 > $O/r2a_ncu_groupby.log 2>&1
timeout 900 $NCU -c 3 -o $O/r2a_prof_motion python scripts/bench_ops.py motion --rows 4e7 --steps 1 --warmup 0 > $O/r2a_ncu_motion.log 2>&1
timeout 900 $NCU -c 12 -o $O/r2a_prof_sort python scripts/bench_ops.py sort --rows 4e7 --steps 1 --warmup 0 > $O/r2a_ncu_sort.log 2>&1
timeout 900 $NCU -c 2 -o $O/r2a_prof_narrow python scripts/dev_q1.py 4e7 narrow > $O/r2a_ncu_narrow.log 2>&1
ls -la $O | tail -30
