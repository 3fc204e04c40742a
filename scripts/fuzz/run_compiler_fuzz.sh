#!/bin/bash
# Build the plan compiler standalone with ASan + UBSan and fuzz it (scripts/fuzz/compiler_fuzz.py).  CPU only.
set -e
cd "$(dirname "$0")/../.."
OUT=$(mktemp -d)
( cd greengage_b200/csrc && g++ -O1 -g -std=c++17 -fPIC -Wall -Wextra -fsanitize=address,undefined -fno-omit-frame-pointer -shared -I . \
    -o "$OUT/libfz.so" gg_compile.cpp ../../scripts/fuzz/compiler_wrap.cpp )
FZ_LIB="$OUT/libfz.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python scripts/fuzz/compiler_fuzz.py
rm -rf "$OUT"
