"""Opcode histogram (weighted by executed warp instructions / stall samples) from `ncu --page source --csv`."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
nvox = float(sys.argv[2]) if len(sys.argv) > 2 else 419.4e6
hdr = rows[1]
i_src, i_ex, i_s, i_thr = (hdr.index(k) for k in ('Source', 'Instructions Executed', '# Samples', 'Avg. Threads Executed'))
data = [r for r in rows[2:] if len(r) == len(hdr) and r[i_ex].isdigit() and r[0].startswith('0x')]
tot = sum(int(r[i_ex]) for r in data); tots = sum(int(r[i_s]) for r in data)
print('sass lines', len(data), 'warp-instr', tot, 'samples', tots, 'lane-instr/voxel %.1f' % (tot * 32 / nvox))
c = Counter(); cs = Counter()
for r in data:
    t = r[i_src].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[i_ex]); cs[op] += int(r[i_s])
for op, n in c.most_common(30):
    print(f"{op:12s} {n / tot * 100:6.2f}% exec  {cs[op] / max(tots,1) * 100:6.2f}% samples")
print('avg active threads %.1f' % (sum(float(r[i_thr]) * int(r[i_ex]) for r in data) / tot))
if len(sys.argv) > 3:
    top = sorted(data, key=lambda r: -int(r[i_s]))[:int(sys.argv[3])]
    for r in top: print(r[i_s].rjust(6), r[i_ex].rjust(10), r[i_src][:100])
