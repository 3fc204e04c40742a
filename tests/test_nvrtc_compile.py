"""The run-time specialised kernel sources compile with NVRTC for sm_100a — on the CPU, with the headers and options
gg_jit.cpp uses (scripts/nvrtc_check.py).  A GPU box whose kernels fail to specialise would fall back to the interpreter
kernels silently; this catches header-registration and syntax problems before that."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_specialised_kernels_compile_with_nvrtc():
    try:
        ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc is not installed")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "nvrtc_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("rc=0") >= 3
