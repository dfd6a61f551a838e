"""Headline metrics of an `ncu --set full` report per kernel + hottest source lines (profiles/ncu_full_*.txt).
usage: python tools/ncu_summary.py <report.ncu-rep> <lib.so> <out.txt> ["header note"]"""
import csv, os, subprocess, sys
rep, lib, out = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); hdr = r[0]; units = r[1]; idx = {n: i for i, n in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__grid_size', 'launch__block_size']
with open(out, "w") as f:
    f.write("# ncu --set full --import-source on --clock-control none, one launch of each sub-step kernel; report %s\n" % rep)
    if note:
        f.write("# " + note + "\n")
    for row in r[2:]:
        f.write("== %s\n" % row[idx['Kernel Name']])
        for n in want:
            if n in idx:
                f.write('   %-90s %16s %s\n' % (n, row[idx[n]], units[idx[n]]))
    here = os.path.dirname(os.path.abspath(__file__))
    for k in ("solve", "dynamics", "collide"):
        f.write("-- hottest source lines: avg_%s_kernel\n" % k)
        f.write(subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, lib, "avg_" + k, "14"], capture_output=True, text=True).stdout)
