#!/usr/bin/env python3
"""Opcode mix of one kernel from an ncu report's source page (SASS view).
usage: ncu_opmix.py report.ncu-rep kernel_regex
Prints executed warp instructions per opcode and the top stall samples per opcode."""
import csv, subprocess, sys, collections, io
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be concatenated: take the first block
hdr = None
ops = collections.Counter(); samples = collections.Counter()
nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1: break
        print(r[1]); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    src = r[hdr.index("Source")].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.rstrip(";")
    base = op.split(".")[0]
    key = base if base not in ("LDG", "LDS", "STS", "STG", "LD", "ST") else op
    ops[key] += int(r[hdr.index("Instructions Executed")])
    samples[key] += int(r[hdr.index("# Samples")])
tot = sum(ops.values()); ts = sum(samples.values())
print("total warp instructions", tot)
for k, v in ops.most_common(40):
    print(f"{k:28s} {v:14d} {100.0*v/tot:6.2f}%   samples {100.0*samples[k]/max(ts,1):6.2f}%")
