"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run on the CPU box).

    python tools/profile_report.py <tag> --launches gpurun_out/launches_X.csv [...] --reps gpurun_out/prof_Y.ncu-rep [...]
"""
import argparse, csv, io, json, re, subprocess, sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
METRICS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
           'sm__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def short(name):
    name = re.sub(r"void |rpx::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name[:90]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        n = short(r[ki])
        us = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    out = [f"| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c} | {us:.1f} | {100*us/tot:.1f} % |")
    return "\n".join(out), tot


def rep_metrics(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        d = OrderedDict(kernel=short(r[hdr.index('Kernel Name')]))
        for m in METRICS:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {rows[1][hdr.index(m)]}"
        out.append(d)
    return out


ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--launches", nargs="*", default=[])
ap.add_argument("--reps", nargs="*", default=[])
ap.add_argument("--note", default="")
a = ap.parse_args()
md = [f"# ncu summary `{a.tag}`", "", a.note, ""]
for p in a.launches:
    table, tot = launches(p)
    md += [f"## launch list `{Path(p).name}` (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares) — total {tot/1e3:.2f} ms", "", table, ""]
    (ROOT / "profiles" / Path(p).name).write_text(Path(p).read_text())
for p in a.reps:
    md += [f"## `ncu --set full` capture `{Path(p).name}`", ""]
    for d in rep_metrics(p):
        md.append(f"### `{d.pop('kernel')}`")
        md += [f"- {k}: {v}" for k, v in d.items()]
        md.append("")
(ROOT / "profiles" / f"{a.tag}.md").write_text("\n".join(md))
print("wrote", ROOT / "profiles" / f"{a.tag}.md")
