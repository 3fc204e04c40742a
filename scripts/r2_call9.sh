#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): the multi-segment tests over the C interconnect (NCCL) and the bench line at N = 2 the way the
# driver launches it.
mkdir -p gpurun_out
O=gpurun_out
T=r2i
nvidia-smi -L | head -4
timeout -s KILL 300 python -m pytest tests/test_gpu_multiseg.py -q -x > $O/${T}_pytest_multiseg.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${T}_pytest_multiseg.log; tail -8 $O/${T}_pytest_multiseg.log
if [ $rc -ge 124 ]; then echo "hang: stopping"; exit 1; fi
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
echo "bench rc=$?" >> $O/${T}_bench_n2.err
tail -c 1800 $O/${T}_bench_n2.json; tail -5 $O/${T}_bench_n2.err
