#!/usr/bin/env python
"""Compile the run-time specialised kernel sources with NVRTC on the CPU (no GPU needed): the same headers and options
gg_jit.cpp hands to nvrtcCompileProgram.  Catches header-registration and syntax problems before a GPU call."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from greengage_b200 import aocs, capi, tpch  # noqa: E402

L = capi.dev_lib()
L.gg_debug_jit_source.argtypes = [C.POINTER(capi.gg_scan), C.POINTER(capi.gg_agg), C.POINTER(capi.gg_exprpool), C.c_int, C.c_int,
                                  C.c_char_p, C.POINTER(C.c_ulonglong), C.c_char_p, C.c_int, C.c_int]
N = C.CDLL("libnvrtc.so.12")
FAKE = ("typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
        "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
        "typedef unsigned long long uintptr_t;\n#define INT32_MIN (-2147483647 - 1)\n#define INT32_MAX 2147483647\n"
        "#define INT64_MIN (-9223372036854775807LL - 1)\n#define INT64_MAX 9223372036854775807LL\n")
rd = lambda *p: open(os.path.join(ROOT, *p)).read()
HDR = [("../../include/gg_plan.h", rd("include", "gg_plan.h")), ("gg_program.h", rd("greengage_b200", "csrc", "gg_program.h")),
       ("gg_device.cuh", rd("greengage_b200", "csrc", "gg_device.cuh")), ("gg_scanagg_kernel.cuh", rd("greengage_b200", "csrc", "gg_scanagg_kernel.cuh")),
       ("stdint.h", FAKE), ("cuda_runtime.h", ""), ("gg_plan.h", rd("include", "gg_plan.h")), ("../../include/gg_aocs.h", rd("include", "gg_aocs.h")),
       ("gg_aocs_decode.h", rd("greengage_b200", "csrc", "gg_aocs_decode.h"))]


def compile_src(src, label):
    prog = C.c_void_p()
    names = (C.c_char_p * len(HDR))(*[n.encode() for n, _ in HDR])
    srcs = (C.c_char_p * len(HDR))(*[s.encode() for _, s in HDR])
    assert N.nvrtcCreateProgram(C.byref(prog), src.encode(), b"gg_jit_scanagg.cu", len(HDR), srcs, names) == 0
    opts = (C.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo")
    rc = N.nvrtcCompileProgram(prog, 3, opts)
    sz = C.c_size_t()
    N.nvrtcGetProgramLogSize(prog, C.byref(sz))
    log = C.create_string_buffer(sz.value + 1)
    N.nvrtcGetProgramLog(prog, log)
    print("%-28s rc=%d %s" % (label, rc, log.value.decode()[:1500] if rc else ""))
    return rc


def source(scan, agg, pool, mode, threads=672, regs=-1):
    buf = C.create_string_buffer(1 << 18)
    h = C.c_ulonglong(0)
    n = L.gg_debug_jit_source(C.byref(scan), C.byref(agg), C.byref(pool), mode, threads, b"", C.byref(h), buf, 1 << 18, regs)
    assert n > 0, L.gg_last_error()
    return buf.value.decode()


bad = 0
scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE)
bad += compile_src(source(scan, agg, pool, 0), "q1 heap priv")
bad += compile_src(source(scan, agg, pool, 1, 256, 0), "q1 heap transposed")
types = [capi.FLOAT8OID] * 4 + [capi.BPCHAROID] * 2 + [capi.DATEOID]
names = dict(quantity=1, extendedprice=2, discount=3, tax=4, returnflag=5, linestatus=6, shipdate=7)
rscan, ragg, rpool = tpch.q1_plan(desc=capi.rows_tupdesc(types, notnull=[1] * 7), cols=names)
bad += compile_src(source(rscan, ragg, rpool, 0), "q1 rows / aocs priv")
sys.exit(1 if bad else 0)
