#!/usr/bin/env python3
"""Aggregate an ncu --import-source capture by source-line ranges of one file.
usage: ncu_ranges.py <rep> <pixels> <file> name:lo-hi ..."""
import csv, io, subprocess, sys, collections
rep, npix, fname = sys.argv[1], float(sys.argv[2]), sys.argv[3]
ranges = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
hdr = None; fpath = ""
acc = collections.OrderedDict((n, [0, 0, 0]) for n, _, _ in ranges); acc["other:" + fname] = [0, 0, 0]; other = collections.Counter()
tot = [0, 0, 0]
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if len(r) == 2 and r[0] == "File Path": fpath = r[1]; continue
    if hdr is None or len(r) < 10 or not r[0]: continue
    try: ie = int(r[hdr["Instructions Executed"]]); sm = int(r[hdr["# Samples"]]); wf = int(r[hdr["L1 Wavefronts Shared"]] or 0)
    except ValueError: continue
    line = int(r[0]); key = None
    if fpath.endswith(fname):
        for n, lo, hi in ranges:
            if lo <= line <= hi: key = n; break
        if key is None: key = "other:" + fname
        a = acc[key]
    else:
        a = acc.setdefault(fpath.split("/")[-1], [0, 0, 0])
    a[0] += ie; a[1] += sm; a[2] += wf
    tot[0] += ie; tot[1] += sm; tot[2] += wf
print("total: %.1f thread-instr/px, %d samples, %.3f wavefronts/px" % (tot[0] * 32 / npix, tot[1], tot[2] / npix))
for k, a in acc.items():
    print("%-28s inst %5.1f%% (%5.1f/px)  samples %5.1f%%  wavefronts %5.1f%%" % (k, 100.0 * a[0] / tot[0], a[0] * 32 / npix, 100.0 * a[1] / tot[1], 100.0 * a[2] / max(tot[2], 1)))
