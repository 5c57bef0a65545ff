#!/usr/bin/env python3
"""Per-source-line view of an ncu capture taken with --import-source on (kernels built with -lineinfo):
instructions executed, stall samples, shared-memory wavefronts per CUDA source line.
usage: ncu_lines.py <report.ncu-rep> [pixels-per-launch]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
npix = float(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
acc = collections.OrderedDict()
cur = None
tot = [0, 0, 0, 0]
for r in rows:
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # two "Source" columns: first CUDA, second SASS
        continue
    if hdr is None or len(r) < 10:
        if len(r) == 2 and r[0] == "File Path": fpath = r[1]
        continue
    line, src = r[0], r[1]
    try:
        ie = int(r[hdr["Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
        wf = int(r[hdr["L1 Wavefronts Shared"]] or 0) if "L1 Wavefronts Shared" in hdr else 0
        wfi = int(r[hdr["L1 Wavefronts Shared Ideal"]] or 0) if "L1 Wavefronts Shared Ideal" in hdr else 0
    except ValueError:
        continue
    if not line:   # SASS row: already counted in its CUDA line
        continue
    key = (fpath.split("/")[-1], line)
    a = acc.setdefault(key, [src.strip()[:100], 0, 0, 0, 0])
    a[1] += ie; a[2] += sm; a[3] += wf; a[4] += wfi
    tot[0] += ie; tot[1] += sm; tot[2] += wf; tot[3] += wfi
print("total warp-instr %d  samples %d  smem wavefronts %d (ideal %d)" % tuple(tot))
if npix: print("per pixel: %.1f thread-instr, %.3f wavefronts" % (tot[0] * 32 / npix, tot[2] / npix))
for (f, line), a in acc.items():
    if a[1] * 200 < tot[0] and a[2] * 200 < tot[1] and a[3] * 200 < max(tot[2], 1): continue
    print("%-14s %4s  inst %5.1f%%  samp %5.1f%%  wf %5.1f%% (x%.2f)  %s" % (f, line, 100.0 * a[1] / tot[0], 100.0 * a[2] / max(tot[1], 1), 100.0 * a[3] / max(tot[2], 1), a[3] / max(a[4], 1), a[0]))
