"""Turns the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarise_profiles.py <launches.csv> <report.ncu-rep> <tag>
"""
import collections
import csv
import json
import subprocess
import sys

launches, rep, tag = sys.argv[1:4]


def short(name):
    return name.split("(")[0].replace("unnamed>::", "").replace("void ", "").split("<")[0].strip()


rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(short(r[ki]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi])
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list, {tag} (workload c3: 1,000,004 bodies, astro theta=1.3 + verlet)", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 108 --csv python bench.py --steps 5 --warmup 3 --skip-extras`",
       "(the first 60 launches - the warm-up steps, which use the global LSD sort before the first host check - are skipped)",
       f"({len(rows)} launches = {len(rows) // 9} steps of 9 kernels; cold-cache, serialised: compare SHARES with bench.py's live `roofline.kernels`)", "",
       "| kernel | launches | total us | us / launch | share |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {v[1] / tot * 100:.1f}% |")
open(f"profiles/{tag}_launches_c3.md", "w").write("\n".join(out) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size", "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio"}
cols = {k: v for k, v in cols.items() if v in hdr}
agg = collections.OrderedDict()
for r in rows[2:]:
    a = agg.setdefault(short(r[hdr.index("Kernel Name")]), {"n": 0, **{k: 0.0 for k in cols}})
    a["n"] += 1
    for k, v in cols.items():
        try:
            a[k] += float(r[hdr.index(v)].replace(",", ""))
        except ValueError:
            pass
out = [f"# ncu --set full summary, {tag} (c3: 1,000,004 bodies, astro theta=1.3 + verlet)", "",
       "`ncu --set full --clock-control none --import-source on -s 51 -c 11 python bench.py --steps 5 --warmup 3 --skip-extras`",
       "(one step; caches flushed between replays, so DRAM traffic is the cold-L2 figure; averages per launch)", "",
       "| kernel | launches | us | DRAM read MB | DRAM write MB | DRAM % peak | SM % peak | issue active % | warps active % | active lanes / instr | regs | grid |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
js = {}
for k, a in agg.items():
    n = a["n"]
    out.append(f"| {k} | {n} | {a['t'] / n:.1f} | {a['rd'] / n:.1f} | {a['wr'] / n:.1f} | {a.get('dram_pct', 0) / n:.1f} | "
               f"{a.get('sm_pct', 0) / n:.1f} | {a.get('issue', 0) / n:.1f} | {a.get('warps', 0) / n:.1f} | "
               f"{a.get('lanes', 0) / n:.1f} | {a.get('regs', 0) / n:.0f} | {a.get('grid', 0) / n:.0f} |")
    js[k] = {"launches_captured": n, "us_per_launch": a["t"] / n, "dram_read_mb_per_launch": a["rd"] / n,
             "dram_write_mb_per_launch": a["wr"] / n}
open(f"profiles/{tag}_ncu_c3_summary.md", "w").write("\n".join(out) + "\n")
json.dump(js, open(f"profiles/{tag}_ncu_c3_kernels.json", "w"), indent=1)
print("\n".join(out[5:]))
