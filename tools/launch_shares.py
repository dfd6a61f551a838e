"""Summarise an `ncu --metrics gpu__time_duration.sum,...` launch list (csv) per kernel. usage: launch_shares.py file.csv n_env"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID')
d = collections.defaultdict(dict)
for r in rows[1:]:
    d[(r[iid], r[ik].split('(')[0][:40])][r[im]] = float(r[iv].replace(',', ''))
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for (i, k), m in d.items():
    for mm, v in m.items(): agg[k][mm].append(v)
tot = sum(sum(v['gpu__time_duration.sum']) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = v['gpu__time_duration.sum']
    extra = ""
    if 'smsp__inst_executed.sum' in v:
        extra = (f" inst/env {sum(v['smsp__inst_executed.sum'])/len(t)/n_env:8.0f} issue {sum(v['smsp__issue_active.avg.pct_of_peak_sustained_active'])/len(t):5.1f}%"
                 f" warps_active {sum(v['sm__warps_active.avg.pct_of_peak_sustained_active'])/len(t):5.1f}%")
    print(f"{k:40s} n={len(t):3d} mean {sum(t)/len(t)/1e6:8.3f} ms share {sum(t)/tot*100:5.1f}%{extra}")
