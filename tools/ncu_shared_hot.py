"""Shared-memory wavefronts per SASS instruction of one kernel from an ncu report (source page):
usage ncu_shared_hot.py report.ncu-rep kernel_regex [top]  -> instructions with the most excess wavefronts."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
items = []
nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1: break
        continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    w, ideal, ex = (int(r[hdr.index(k)] or 0) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive"))
    if w: items.append((ex, w, ideal, int(r[hdr.index("Instructions Executed")]), r[hdr.index("Source")].strip()[:90]))
tw = sum(i[1] for i in items); te = sum(i[0] for i in items)
print(f"shared wavefronts {tw}  excessive {te} ({100*te/max(tw,1):.1f} %)")
for ex, w, ideal, n, src in sorted(items, reverse=True)[:top]:
    print(f"excess {ex:10d}  wavefronts {w:10d}  ideal {ideal:10d}  executed {n:9d}  {src}")
