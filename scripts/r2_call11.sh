#!/bin/bash
# Round 2, GPU call 11 (the last minute of the budget): one `--set full` capture of the headline scan kernel as it stands at the
# end of the round (Q1 over 2x10^7-row lineitem-wide: 3.4 GB >> L2, same kernel, same launch configuration as the bench), then
# the launch list of the same command.
mkdir -p gpurun_out
O=gpurun_out
CMD="python scripts/ab_scan.py 2e7"
timeout -s KILL 45 ncu --set full --clock-control none --import-source on -k regex:gg_jit_scanagg --launch-skip 3 -c 1 -f -o $O/r2k_prof_scanagg $CMD > $O/r2k_ncu_scanagg.log 2>&1
tail -3 $O/r2k_ncu_scanagg.log
timeout -s KILL 30 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2k_launches.csv $CMD > $O/r2k_ncu_launches.log 2>&1
grep -c gg_ $O/r2k_launches.csv
