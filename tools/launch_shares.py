"""Summarise an `ncu --metrics gpu__time_duration.sum,...` launch list (csv) per kernel; optionally write the per-step
totals bench.py reads (profiles/traffic.json). usage: launch_shares.py file.csv n_env [steps_in_list] [traffic.json]"""
import csv, collections, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID'); iu = hdr.index('Metric Unit')
d = collections.defaultdict(dict)
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    u = r[iu].lower()
    if r[im].startswith('dram__bytes'):
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if r[im] == 'gpu__time_duration.sum':
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
    d[(r[iid], r[ik].split('(')[0][:40])][r[im]] = v
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for (i, k), m in d.items():
    for mm, v in m.items(): agg[k][mm].append(v)
# launches of each of our kernels in one env-step of one (half-)batch: a list cut at an arbitrary launch is normalised by these
PER_STEP = {"avg_dynsolve": 5, "avg_solve": 5, "avg_dynamics": 5, "avg_collide": 5, "avg_narrow": 5, "avg_epilogue": 1, "avg_prologue": 1}
def per_step(k):
    return next((c for n, c in PER_STEP.items() if n in k), 0)
tot = sum(sum(v['gpu__time_duration.sum']) for v in agg.values())
tot_inst = tot_dram = 0.0
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = v['gpu__time_duration.sum']
    extra = ""
    if 'smsp__inst_executed.sum' in v:
        tot_inst += sum(v['smsp__inst_executed.sum'])
        extra = (f" warp-inst/env/launch {sum(v['smsp__inst_executed.sum'])/len(t)/n_env:8.0f} issue {sum(v['smsp__issue_active.avg.pct_of_peak_sustained_active'])/len(t):5.1f}%"
                 f" warps_active {sum(v['sm__warps_active.avg.pct_of_peak_sustained_active'])/len(t):5.1f}%")
    if 'dram__bytes_read.sum' in v:
        db = sum(v['dram__bytes_read.sum']) + sum(v['dram__bytes_write.sum'])
        tot_dram += db
        extra += f" dram/env/launch {db/len(t)/n_env:7.0f} B"
    if 'smsp__thread_inst_executed_per_inst_executed.ratio' in v:
        extra += f" lanes/inst {sum(v['smsp__thread_inst_executed_per_inst_executed.ratio'])/len(t):4.1f}"
    print(f"{k:40s} n={len(t):3d} mean {sum(t)/len(t)/1e6:8.3f} ms share {sum(t)/tot*100:5.1f}%{extra}")
step_inst = sum(per_step(k) * sum(v['smsp__inst_executed.sum']) / len(v['smsp__inst_executed.sum']) for k, v in agg.items() if 'smsp__inst_executed.sum' in v)
step_dram = sum(per_step(k) * (sum(v['dram__bytes_read.sum']) + sum(v['dram__bytes_write.sum'])) / len(v['dram__bytes_read.sum']) for k, v in agg.items() if 'dram__bytes_read.sum' in v)
step_time = sum(per_step(k) * sum(v['gpu__time_duration.sum']) / len(v['gpu__time_duration.sum']) for k, v in agg.items())
print(f"one env-step (per-kernel means x launches per step): {step_inst/n_env:.0f} warp instructions / env, {step_dram/n_env:.0f} DRAM bytes / env, serialised kernel time {step_time/1e6:.3f} ms for {n_env} envs")
tot_inst, tot_dram, n_steps, tot = step_inst, step_dram, 1.0, step_time
print(f"per env-step: {tot_inst/n_steps/n_env:.0f} warp instructions, {tot_dram/n_steps/n_env:.0f} DRAM bytes (read+write), serialised kernel time {tot/n_steps/1e6:.3f} ms")
if len(sys.argv) > 4:
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_hash
    json.dump({"source": sys.argv[1], "n_env": n_env, "kernel_hash": kernel_source_hash(), "dram_bytes_per_step_per_env": tot_dram / n_steps / n_env,
               "warp_inst_per_env_step": tot_inst / n_steps / n_env,
               "issue_slots": {"warp_inst_per_env_step": tot_inst / n_steps / n_env, "peak_warp_inst_per_s": 148 * 4 * 1.965e9,
                               "note": "issue-slot roof = 148 SMs x 4 schedulers x 1.965 GHz; frac = env-steps/s x warp_inst_per_env_step / peak"}},
              open(sys.argv[4], "w"), indent=1)
