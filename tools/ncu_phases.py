"""Development helper: aggregate ncu source-page stall samples per barrier-delimited phase of a kernel.
usage: ncu_phases.py report.ncu-rep kernel_regex"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
phase, agg, tot = 0, {}, 0
seen = set()
for r in rows[hi + 1:]:
    if len(r) <= max(ia, isamp, iex):
        continue
    if r[0] in seen:  # report lists each launch once per matching ID; keep the first
        break
    seen.add(r[0])
    try:
        s, ex = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    toks = r[ia].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    a = agg.setdefault(phase, [0, 0, {}])
    a[0] += s
    a[1] += ex
    a[2][op] = a[2].get(op, 0) + s
    tot += s
    if "BAR.SYNC" in r[ia]:
        phase += 1
for p, (s, ex, ops) in agg.items():
    top = sorted(ops.items(), key=lambda x: -x[1])[:6]
    print(f"phase {p}: samples {100 * s / max(tot, 1):5.1f}%  warp-instr {ex / 1e6:8.1f}M  top {top}")
