#!/bin/bash
# Round 2, GPU call 10 (2 GPUs, short): where the Redistribute-HashJoin test over two NCCL segments gets stuck (collective and
# executor-node traces on stderr), then the bench line at N = 2 the way the driver launches it, on a fifth of the rows (what is
# checked here is that the N > 1 path runs and agrees with the oracle, not its speed), the Redistribute-HashJoin in its
# killable child processes.
mkdir -p gpurun_out
O=gpurun_out
T=r2j
GGB200_IC_TRACE=1 GGB200_EXEC_TRACE=1 timeout -s KILL 55 python -m pytest "tests/test_gpu_multiseg.py::test_redistribute_hashjoin_over_nccl_segments[q3ish]" -q -x -s > $O/${T}_q3ish.log 2>&1
echo "q3ish rc=$?" >> $O/${T}_q3ish.log; tail -c 3000 $O/${T}_q3ish.log
GGB200_IC_TRACE=1 GGB200_EXEC_TRACE=1 GGB200_RJOIN_TIMEOUT=45 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --rows 2e7 --rjoin-rows 4e7 --secondary rjoin > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
echo "bench rc=$?" >> $O/${T}_bench_n2.err
cut -c1-1500 $O/${T}_bench_n2.json; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2j_bench_n2.json"))
    print("SECONDARY", json.dumps(d.get("secondary"))[:4000])
except Exception as e:
    print("no line", e)
PY
grep -v "^\[ic seg\|^\[exec seg" $O/${T}_bench_n2.err | tail -8
