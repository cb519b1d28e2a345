"""Per-instruction memory cost of one profiled kernel, from the SASS page of an `ncu --set full --import-source on` report:
shared-memory wavefronts and L1 tag requests of every LDS / STS / LDG / STG, divided by a unit count (e.g. frames), plus the
stall samples by region.   python tools/ncu_memops.py <report.ncu-rep> <units> [> profiles/...]"""
import csv
import io
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:100])
hdr, data = rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
num = lambda r, h: int(r[col[h]] or 0)
tot_s = tot_g = 0
print("%5s  %-62s %8s %10s %10s %10s" % ("#", "instruction", "exec/u", "smem wf/u", "ideal/u", "L1 tags/u"))
for n, r in enumerate(data):
    src = r[col["Source"]].strip()
    ex = num(r, "Instructions Executed")
    if ex == 0 or not any(k in src for k in ("LDS", "STS", "LDG", "STG", "LDL", "STL")):
        continue
    w, wi, tag = num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Ideal"), num(r, "L1 Tag Requests Global")
    tot_s += w
    tot_g += tag
    if (w + tag) / units >= 0.5:
        print("%5d  %-62s %8.2f %10.1f %10.1f %10.1f" % (n, src[:62], ex / units, w / units, wi / units, tag / units))
print("total per unit: %.0f shared-memory wavefronts, %.0f L1 tag requests (rows below 0.5 per unit not listed)" % (tot_s / units, tot_g / units))
samples = sum(num(r, "# Samples") for r in data)
for nm in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_barrier", "stall_selected", "stall_not_selected"):
    print("%-20s %5.1f %% of the stall samples" % (nm, 100.0 * sum(num(r, nm) for r in data) / max(samples, 1)))
