"""Summarise an ncu report exported with
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
usage: python profiles/analyze.py raw.csv sass.csv [top_n]"""
import csv, sys
from collections import Counter
raw, sass = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(raw)))
hdr = rows[0]
for w in ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
          'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
          'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
          'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum']:
    if w in hdr:
        print('%-70s %s %s' % (w, rows[2][hdr.index(w)], rows[1][hdr.index(w)]))
rows = list(csv.reader(open(sass)))
hdr = rows[1]
si, ie, src = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[si]) for r in data)
print('sass instructions', len(data), 'samples', tot, 'warp-instr executed', sum(int(r[ie]) for r in data))
for name in ['stall_long_sb', 'stall_short_sb', 'stall_no_inst', 'stall_wait', 'stall_math', 'stall_branch_resolving',
             'stall_barrier', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_lg', 'stall_mio', 'stall_membar']:
    i = hdr.index(name)
    v = sum(int(r[i]) for r in data)
    print('  %-24s %8d  %5.1f%%' % (name, v, 100.0 * v / max(tot, 1)))
c, cs = Counter(), Counter()
for r in data:
    t = r[src].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[ie]); cs[op] += int(r[si])
print('opcode: executed, samples')
for k, v in c.most_common(18):
    print('  %-8s %12d %8d' % (k, v, cs[k]))
print('top stalled instructions: samples, executed, avg threads, sass')
at = hdr.index('Avg. Threads Executed')
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    print('  %7s %10s %4s  %s' % (r[si], r[ie], r[at], r[src][:80]))
