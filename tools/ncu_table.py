"""One table of the headline ncu metrics of every kernel in a report: python tools/ncu_table.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('gpu__time_duration.sum', 'us'),
        ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'), ('smsp__inst_executed.sum', 'warp inst'),
        ('sm__inst_executed.avg.per_cycle_active', 'ipc'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ %'), ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu %'), ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma %'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu %'), ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu wavefronts %'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bank conflicts'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected')]
units = rows[1]
for name, label in want:
    if name in h:
        i = h.index(name)
        vals = []
        for r in rows[2:]:
            v = r[i]
            try:
                f = float(v)
                v = ('%.0f' % f) if abs(f) >= 100 else ('%.2f' % f)
            except ValueError:
                v = v.split('(')[0]
            vals.append(v)
        print('%-20s %-8s %s' % (label, units[i], '  '.join('%12s' % v for v in vals)))
