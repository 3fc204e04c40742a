#!/bin/bash
# What the next session with GPU minutes should run first, cheapest first (each block is one gpurun call; every step has its own
# timeout — a step that hangs must never take the call's whole limit with it, see profiles/README.md "third session").
#
#   1 GPU, ~6 min:   gpurun --timeout 900 -- 'bash scripts/next_gpu_calls.sh one'
#   2 GPUs, ~4 min:  gpurun --gpus 2 --timeout 400 -- 'bash scripts/next_gpu_calls.sh two'
#
# `two` is the one that matters: the Redistribute-HashJoin's row exchange was fixed after the last 2-GPU minute was spent
# (gg_ic_allgather_u64 used the count matrix as scratch) and has not run on hardware since.
mkdir -p gpurun_out
O=gpurun_out
case "$1" in
one)
  timeout -s KILL 700 python -m pytest tests -m gpu -q -x > $O/next_pytest.log 2>&1; echo "pytest rc=$?" >> $O/next_pytest.log; tail -5 $O/next_pytest.log
  timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > $O/next_bench_n1.json 2> $O/next_bench_n1.err; echo "bench rc=$?"; tail -c 600 $O/next_bench_n1.json
  ;;
two)
  GGB200_IC_TRACE=1 GGB200_EXEC_TRACE=1 timeout -s KILL 120 python -m pytest tests/test_gpu_multiseg.py -q -x -s > $O/next_multiseg.log 2>&1
  echo "multiseg rc=$?" >> $O/next_multiseg.log; grep -v "^\[ic seg\|^\[exec seg" $O/next_multiseg.log | tail -8
  GGB200_RJOIN_TIMEOUT=90 timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 10 --warmup 3 --secondary rjoin > $O/next_bench_n2.json 2> $O/next_bench_n2.err
  echo "bench rc=$?"; grep "^{" $O/next_bench_n2.json | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d.get('secondary'))[:1500])"
  # if the point-to-point exchange is still stuck: the same with GGB200_IC_ROWS=auto (proof + all-gather fallback), then =allgather
  ;;
*) echo "usage: $0 one|two"; exit 2;;
esac
