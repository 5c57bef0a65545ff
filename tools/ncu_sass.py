#!/usr/bin/env python3
"""SASS-level view of an ncu capture (--import-source on, -lineinfo): executed warp-instructions per opcode and the
hottest SASS instructions, optionally restricted to a range of CUDA source lines of one file.
usage: ncu_sass.py <report.ncu-rep> [pixels-per-launch] [file:first-last]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
npix = float(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = None
if len(sys.argv) > 3:
    f, r = sys.argv[3].split(":")
    a, b = r.split("-")
    rng = (f, int(a), int(b))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fpath, cur = None, "", None
ops = collections.Counter()
tot = 0
listing = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10:
        if len(r) == 2 and r[0] == "File Path": fpath = r[1].split("/")[-1]
        continue
    if r[0]:   # CUDA line row
        cur = (fpath, int(r[0]))
        continue
    try:
        ie = int(r[hdr["Instructions Executed"]])
    except ValueError:
        continue
    if rng and not (cur and cur[0] == rng[0] and rng[1] <= cur[1] <= rng[2]): continue
    sass = r[3].strip()
    if not sass or sass == "...": continue
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    ops[op.split(".")[0]] += ie
    tot += ie
    listing.append((cur, sass, ie))
print("warp-instructions %d%s" % (tot, "  = %.2f thread-instr/px" % (tot * 32 / npix) if npix else ""))
for op, n in ops.most_common(24):
    print("  %-10s %5.1f%%%s" % (op, 100.0 * n / tot, "  %.2f/px" % (n * 32 / npix) if npix else ""))
if rng:
    for cur, sass, ie in listing:
        print("%4d %9d  %s" % (cur[1], ie, sass[:90]))
