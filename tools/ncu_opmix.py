#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an .ncu-rep captured with --import-source on.
usage: tools/ncu_opmix.py X.ncu-rep [n_regions]  -> opcode histogram (warp-level executed counts), sample share by address region"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ops = collections.Counter(); samples = collections.Counter(); tot = 0; stot = 0
base = int(rows[0]["Address"], 16)
for r in rows:
    src = r["Source"].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    n = int(r["Instructions Executed"] or 0)
    s = int(r["# Samples"] or 0)
    fam = op.split(".")[0]
    if fam in ("IMAD",) and ".MOV" in op: fam = "IMAD.MOV"
    ops[fam] += n; samples[fam] += s; tot += n; stot += s
print(f"total warp-level instructions executed: {tot}   samples: {stot}")
for k, v in ops.most_common(32):
    print(f"  {k:12s} {v:12d} {100.0*v/tot:6.2f}%   samples {100.0*samples[k]/max(1,stot):6.2f}%")
# regions between barriers
reg = []; cur = [0, 0, base]
for r in rows:
    cur[0] += int(r["Instructions Executed"] or 0); cur[1] += int(r["# Samples"] or 0)
    if "BAR.SYNC" in r["Source"] or "EXIT" in r["Source"]:
        reg.append((cur[2] - base, int(r["Address"], 16) - base, cur[0], cur[1])); cur = [0, 0, int(r["Address"], 16) + 16]
print("regions (between barriers): start end  inst%  samples%")
for a, b, n, s in reg:
    print(f"  {a:#07x}-{b:#07x}  {100.0*n/tot:6.2f}%  {100.0*s/max(1,stot):6.2f}%")
stall = collections.Counter()
for r in rows:
    for k, v in r.items():
        if k and k.startswith("stall_") and "Not Issued" not in k and v:
            stall[k] += int(v)
print("stall samples:", ", ".join(f"{k[6:]}={v}" for k, v in stall.most_common(10)))
