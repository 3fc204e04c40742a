#!/bin/bash
# A/B builds of libggb200.so from the same sources with extra compiler switches (a -D that an experiment's #ifdef reads, a different
# -maxrregcount, ...) into build/ab_<name>/libggb200.so; measured with
#   GGB200_DEVLIB=build/ab_<name>/libggb200.so python scripts/ab_scan.py 1e8
# (profiles/r2g_ab_scan.jsonl was made this way: the snapshot rule compiled out, teams on the slots' barriers; the switches went
# with the experiment — delete the build/ab_* directories before a gpurun call, they travel with the snapshot)
# usage: scripts/ab_build.sh NAME -DFLAG [...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/ab_$name
mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets -I build"
objs=""
for src in greengage_b200/csrc/*.cu greengage_b200/csrc/*.cpp greengage_b200/csrc/plans/gg_plan_cache.cu; do
  o=$out/$(basename ${src%.*}).o
  $NV "$@" -c $src -o $o &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -Wno-deprecated-gpu-targets -o $out/libggb200.so $objs -ldl
rm -f $out/*.o
ls -la $out
