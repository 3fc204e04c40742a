"""In-tree build of the product libraries (no JIT cache: the .so files travel with the repo snapshot).

  libggb200.so  hand-written CUDA for sm_100a + the C-ABI (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)
  libgghost.so  host C (gcc): synthetic loader, Motion routing
  libggexec.so  host C (gcc): the executor-node surface (ExecInitNode/ExecProcNode/ExecEndNode) over libggb200.so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BUILD = os.path.join(ROOT, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"]
CU_SOURCES = ["gg_abi.cu", "gg_scanagg.cu", "gg_compile.cpp"]
HOST_SOURCES = ["gg_synth.c"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _all_headers():
    hs = []
    for d in (CSRC, HOST, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                hs.append(os.path.join(d, f))
    return hs


def embed_sources():
    """The kernel headers as string literals, for run-time plan specialisation (gg_jit.cpp)."""
    files = [("gg_plan_h", os.path.join(ROOT, "include", "gg_plan.h")), ("gg_program_h", os.path.join(CSRC, "gg_program.h")),
             ("gg_device_cuh", os.path.join(CSRC, "gg_device.cuh")), ("gg_scanagg_kernel_cuh", os.path.join(CSRC, "gg_scanagg_kernel.cuh")),
             ("gg_aocs_h", os.path.join(ROOT, "include", "gg_aocs.h")), ("gg_aocs_decode_h", os.path.join(CSRC, "gg_aocs_decode.h"))]
    out = os.path.join(BUILD, "gg_jit_sources.inc")
    if not _newer(out, [f for _, f in files]):
        return out
    with open(out, "w") as fo:
        for name, path in files:
            text = open(path).read()
            fo.write("static const char *const kSrc_%s =\n" % name)
            for i in range(0, len(text), 8000):        # string literals are limited in length; concatenate pieces
                fo.write('R"GGSRC(%s)GGSRC"\n' % text[i:i + 8000])
            fo.write(";\n")
    return out


def build_device(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    embed_sources()
    objs = []
    hdrs = _all_headers() + [os.path.join(BUILD, "gg_jit_sources.inc")]
    sources = [s for s in os.listdir(CSRC) if s.endswith((".cu", ".cpp"))] + [os.path.join("plans", "gg_plan_cache.cu")]
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        if force or _newer(obj, [path] + hdrs):
            cmd = [NVCC] + NVFLAGS + ["-I", BUILD] + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            subprocess.check_call(cmd)
        objs.append(obj)
    out = os.path.join(HERE, "libggb200.so")
    if force or _newer(out, objs):
        # link to a temporary name and rename: a repo snapshot taken meanwhile never sees a half-written library
        subprocess.check_call([NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", out + ".tmp"] + objs + ["-ldl"])
        os.replace(out + ".tmp", out)
    return out


def build_exec(force=False):
    """libggexec.so: the executor-node surface (host C) above the C-ABI of libggb200.so"""
    out = os.path.join(HERE, "libggexec.so")
    srcs = [os.path.join(HOST, "gg_executor.c"), os.path.join(HOST, "gg_motion_host.c"), os.path.join(HOST, "gg_tupser.c")]
    if force or _newer(out, srcs + _all_headers() + [os.path.join(HERE, "libggb200.so")]):
        subprocess.check_call(["gcc", "-O2", "-g", "-fPIC", "-Wall", "-Wextra", "-std=gnu11", "-shared", "-o", out + ".tmp"] + srcs +
                              ["-L", HERE, "-lggb200", "-Wl,-rpath,$ORIGIN"])
        os.replace(out + ".tmp", out)
    return out


def build_host(force=False):
    out = os.path.join(HERE, "libgghost.so")
    srcs = [os.path.join(HOST, s) for s in os.listdir(HOST) if s.endswith(".c") and s != "gg_executor.c"]
    if force or _newer(out, srcs + _all_headers()):
        subprocess.check_call(["gcc", "-O2", "-g", "-fPIC", "-Wall", "-ffp-contract=off", "-pthread", "-shared",
                               "-o", out + ".tmp"] + srcs + ["-lm"])
        os.replace(out + ".tmp", out)
    return out


def build_tools(force=False):
    """build/gather_peak: the random-access rates of this GPU (scripts/gather_peak.cu) that the join / hash-aggregate rooflines
    of bench.py are held against — a measurement tool, not part of the libraries"""
    src = os.path.join(ROOT, "scripts", "gather_peak.cu")
    out = os.path.join(BUILD, "gather_peak")
    os.makedirs(BUILD, exist_ok=True)
    if force or _newer(out, [src]):
        subprocess.check_call([NVCC, "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Wno-deprecated-gpu-targets",
                               "-o", out + ".tmp", src])
        os.replace(out + ".tmp", out)
    return out


def build_all(verbose=False, force=False):
    res = build_device(verbose, force), build_host(force), build_exec(force)
    build_tools(force)
    return res


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
