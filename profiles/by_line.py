"""Attribute ncu SASS-level counts to CUDA source lines.
usage: python profiles/by_line.py <nvdisasm --print-line-info output> <kernel mangled name> <sass.csv from ncu> [top_n]
The ncu source page (csv) has no per-line metrics for CUDA source, so the SASS rows are zipped, in order,
with the disassembly of the same cubin (which carries //## File "..", line N markers)."""
import csv, re, sys
from collections import defaultdict
dis, kname, sass = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = None
for i, l in enumerate(lines):
    if l.strip().startswith('.section') and ('.text.' + kname) in l:
        start = i
        break
assert start is not None, 'kernel section not found'
cur = ('?', 0)
seq = []
for l in lines[start + 1:]:
    if l.strip().startswith('.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', l):
        seq.append(cur)
rows = list(csv.reader(open(sass)))
hdr = rows[1]
si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
print('disasm instructions', len(seq), 'ncu rows', len(data))
n = min(len(seq), len(data))
ex, sm = defaultdict(int), defaultdict(int)
for k in range(n):
    ex[seq[k]] += int(data[k][ie]); sm[seq[k]] += int(data[k][si])
te, ts = sum(ex.values()), sum(sm.values())
src = {}
print('%-22s %12s %6s %8s %6s' % ('file:line', 'warp-instr', '%', 'samples', '%'))
for key in sorted(ex, key=lambda k: -sm[k])[:topn]:
    f, ln = key
    if f not in src:
        try:
            src[f] = open('/root/repo/bipartitesbm-mcmc_b200/csrc/' + f).read().split('\n')
        except Exception:
            src[f] = []
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ''
    print('%-22s %12d %5.1f%% %8d %5.1f%%  %s' % ('%s:%d' % key, ex[key], 100.0 * ex[key] / te, sm[key], 100.0 * sm[key] / ts, text))
