"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.

usage: ncu_lines.py dump.csv [top_n] [file_substring]
Prints, per source file, the lines with the most stall samples: share of executed warp instructions, share of
samples, and the dominant stall reasons."""
import collections
import csv
import sys


def f(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ''
cur_file, hdr = '', None
lines = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
        continue
    key = (cur_file, int(r[0]))
    a = lines.setdefault(key, [0.0, 0.0, collections.Counter(), r[1][:110]])
    a[0] += f(r[ci])
    a[1] += f(r[si])
    for h, i in stall.items():
        a[2][h] += f(r[i])
ti = sum(v[0] for v in lines.values())
ts = sum(v[1] for v in lines.values())
print('total warp instructions %.4g, samples %d' % (ti, ts))
sel = [(k, v) for k, v in lines.items() if only in k[0]]
for (fn, ln), v in sorted(sel, key=lambda kv: -kv[1][1])[:top_n]:
    top = ', '.join('%s %.0f%%' % (h[6:], 100 * c / max(v[1], 1)) for h, c in v[2].most_common(3))
    print('%-18s %5d %5.1f%% ins %5.1f%% smp | %-34s | %s' % (fn[:18], ln, 100 * v[0] / ti, 100 * v[1] / ts, top, v[3].strip()[:80]))
