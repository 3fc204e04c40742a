#!/bin/bash
# Build the AOCS loader with ASan + UBSan and fuzz it with corrupted / truncated column files (scripts/fuzz/aocs_fuzz.py).
set -e
cd "$(dirname "$0")/../.."
OUT=$(mktemp -d)
gcc -O1 -g -fPIC -Wall -Wextra -fsanitize=address,undefined -fno-omit-frame-pointer -shared -o "$OUT/libfz.so" greengage_b200/host/gg_aocs_host.c tests/aocs_decode_harness.c
FZ_LIB="$OUT/libfz.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python scripts/fuzz/aocs_fuzz.py
rm -rf "$OUT"
