#!/usr/bin/env python
"""Per-kernel SHA-1 of the SASS in greengage_b200/libggb200.so (cuobjdump -sass; no GPU needed).  A host-only change must
leave every line of profiles/*_sass_fingerprint.txt unchanged: that is how round 1's late host-side fixes (plan validation,
argument checks, the separate AOCS translation unit) were shown not to touch the kernels that had been validated on the GPU.
  python scripts/sass_fingerprint.py                      print
  python scripts/sass_fingerprint.py profiles/X.txt       compare with a stored fingerprint"""
import hashlib, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = ("identifier", "arch =", "code version", "host =", "compile_size", "producer")


def fingerprint():
    txt = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-sass", os.path.join(ROOT, "greengage_b200", "libggb200.so")]).decode()
    out, cur, buf = {}, None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if cur:
                out[cur] = hashlib.sha1("\n".join(buf).encode()).hexdigest()
            cur, buf = m.group(1), []
        elif cur and ln.strip() and "Fatbin" not in ln and not ln.startswith(SKIP):
            buf.append(ln)
    if cur:
        out[cur] = hashlib.sha1("\n".join(buf).encode()).hexdigest()
    return out


if __name__ == "__main__":
    fp = fingerprint()
    if len(sys.argv) > 1:
        old = dict(ln.split() for ln in open(sys.argv[1]) if ln.strip() and not ln.startswith("#"))
        changed = sorted(k for k in old if k in fp and fp[k] != old[k])
        print("changed:", changed, "\nremoved:", sorted(k for k in old if k not in fp), "\nadded:", sorted(k for k in fp if k not in old))
        sys.exit(1 if changed else 0)
    for k in sorted(fp):
        print(fp[k], k) if False else print(k, fp[k])
