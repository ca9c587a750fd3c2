"""Per-source-line executed instructions of the kernels in an .ncu-rep (captured with --import-source on):
python tools/ncu_lines.py rep [kernel-substring] [divisor]   -- counts are warp instructions / divisor."""
import csv, subprocess, sys
import os
rep = os.path.abspath(sys.argv[1])
sub = sys.argv[2] if len(sys.argv) > 2 else ""
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True, cwd='/tmp').stdout
rows = list(csv.reader(raw.splitlines()))
kernels, cur, pending = [], None, None
for r in rows:
    if r and r[0] == 'File Path':
        pending = r[1]
    elif r and r[0] == 'Function Name':
        if cur is None or cur['name'] != r[1]:
            cur = {'name': r[1], 'files': []}
            kernels.append(cur)
        cur['files'].append({'file': pending, 'rows': []})
    elif cur is not None and cur['files']:
        cur['files'][-1]['rows'].append(r)
for k in kernels:
    if sub not in k['name']:
        continue
    print('====', k['name'][:110])
    tot = 0
    for f in k['files']:
        hdr = f['rows'][0]
        ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isamp = hdr.index('# Samples')
        lines = [(r[0], r[1], int(r[ie]), int(r[it]), int(r[isamp])) for r in f['rows'][1:] if r[0].isdigit() and r[ie].isdigit()]
        s = sum(l[2] for l in lines)
        tot += s
        if s == 0:
            continue
        print('--', f['file'], f'{s / div:.2f}')
        for ln, src, e, t, sm in lines:
            if e:
                print(f'{ln:>5s} {e / div:8.2f} {t / max(e, 1):5.1f} {sm:6d}  {src[:110]}')
    print('total', tot / div)
