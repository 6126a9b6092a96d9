"""Per-source-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K`.

  python tools/ncu_lines.py <both.csv> [min_share]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr = rows[2]
iL, iS = 0, 1
iN, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = collections.OrderedDict()


def num(x):
    return int(x) if x.isdigit() else 0


for r in rows[3:]:
    if len(r) <= iT or not r[iL].isdigit():
        continue
    a = agg.setdefault(int(r[iL]), [r[iS], 0, 0, 0])
    a[1] += num(r[iN])
    a[2] += num(r[iI])
    a[3] += num(r[iT])
ts, ti = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
print(f"samples {ts}  warp-instructions {ti}")
for ln, a in sorted(agg.items()):
    if a[1] > ts * thr or a[2] > ti * thr:
        lanes = a[3] / a[2] if a[2] else 0
        print(f"{ln:5d} smp {100 * a[1] / ts:5.1f}%  ins {100 * a[2] / ti:5.1f}%  lanes {lanes:4.1f}  {a[0].strip()[:100]}")
