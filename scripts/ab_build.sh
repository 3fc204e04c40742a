#!/bin/bash
# A/B builds of libggb200.so from the same sources with one -D switch each, into build/ab_<name>/libggb200.so; measured with
#   GGB200_DEVLIB=build/ab_<name>/libggb200.so python scripts/sweep_teams.py 1e8 wide default
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
