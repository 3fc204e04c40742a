"""Join an ncu SASS-level source export with nvdisasm line info: executed warp-instructions per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
rows = rows[:hdr_i + 1] + [r for r in rows[hdr_i + 1:] if r and r[0] != "Address"]
ci = hdr.index("Instructions Executed")
si = hdr.index("# Samples") if "# Samples" in hdr else None
sass_counts = [(r[1].strip(), int(r[ci] or 0), int(r[si] or 0) if si is not None else 0) for r in rows[hdr_i + 1:] if len(r) > ci]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    if f.endswith(".cubin") and "sm_100" in f:
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur = None
        infunc = False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                infunc = m.group(1) == kname or m.group(1).endswith("." + kname)
                continue
            if ln.startswith("//----") and ".text." in ln:
                infunc = False
            if not infunc:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
            if m:
                lines.append((cur, m.group(1).strip()))
if len(lines) != len(sass_counts):
    print("warning: %d disasm instrs vs %d profiled" % (len(lines), len(sass_counts)))
agg = collections.Counter()
samp = collections.Counter()
tot = 0
for (loc, _), (_, n, s) in zip(lines, sass_counts):
    agg[loc] += n
    samp[loc] += s
    tot += n
stot = sum(samp.values()) or 1
print("total warp-instructions executed: %d" % tot)
srcs = {}
for loc, n in agg.most_common(top):
    if loc is None:
        print("%6.2f%% %6.2f%%s  <none>" % (100.0 * n / tot, 100.0 * samp[loc] / stot)); continue
    f, l = loc
    if f not in srcs:
        for d in ("greengage_b200/csrc",):
            p = os.path.join(d, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip() if l - 1 < len(srcs[f]) else ""
    print("%6.2f%% %6.2f%%s  %s:%d  %s" % (100.0 * n / tot, 100.0 * samp[loc] / stot, f, l, text[:110]))
