#!/bin/bash
# Build the executor host code + the oracle-backed stand-in with ASan + UBSan and fuzz GgExecInitNode / ProcNode with
# malformed plan trees (scripts/fuzz/executor_fuzz.py).  CPU only.
set -e
cd "$(dirname "$0")/../.."
OUT=$(mktemp -d)
SAN="-O1 -g -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer"
for f in greengage_b200/host/gg_executor.c greengage_b200/host/gg_motion_host.c greengage_b200/host/gg_tupser.c tests/mock/ggb200_mock.c; do gcc $SAN -std=gnu11 -c $f -o "$OUT/$(basename $f).o"; done
for f in tests/mock/compile_glue.cpp greengage_b200/csrc/gg_compile.cpp; do g++ $SAN -std=c++17 -c $f -o "$OUT/$(basename $f).o"; done
g++ -shared -fsanitize=address,undefined -o "$OUT/libfz.so" "$OUT"/*.o -L oracle -lggoracle -Wl,-rpath,"$PWD/oracle" -lm
FZ_LIB="$OUT/libfz.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python scripts/fuzz/executor_fuzz.py
rm -rf "$OUT"
