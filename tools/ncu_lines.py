"""Development helper: warp instructions executed and stall samples per CUDA source line of one kernel.
usage: ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, text, files_seen, ln = None, None, {}, {}, set(), None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if cur in files_seen:  # the report repeats the listing for every captured launch: keep the first
            break
        files_seen.add(cur)
        ln = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        isa, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        ish, ishi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr is None or len(r) <= iex:
        continue
    if r[0] != "":
        try:
            ln = int(r[0])
            text[(cur, ln)] = r[1].strip()
        except ValueError:
            ln = None
        continue
    if ln is None or r[2] in ("", "..."):
        continue

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    a = agg.setdefault((cur, ln), [0, 0, 0, 0, 0])
    a[0] += num(r[isa])
    a[1] += num(r[iex])
    a[2] += 1
    a[3] += num(r[ish])
    a[4] += num(r[ishi])
tot_s = sum(a[0] for a in agg.values()) or 1
tot_e = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, warp instructions {tot_e / 1e6:.1f} M")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[4]) if len(sys.argv) > 4 else 1])[:top]:
    print(f"{f}:{ln:4d}  instr {a[1] / 1e6:7.1f}M ({100 * a[1] / tot_e:4.1f}%)  samples {100 * a[0] / tot_s:4.1f}%  sass {a[2]:4d}  "
          f"smem wf {a[3] / 1e6:6.1f}M (ideal {a[4] / 1e6:6.1f}M)  | {text.get((f, ln), '')[:90]}")
