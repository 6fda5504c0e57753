"""Share of executed warp instructions and stall samples per code region of slode_fixed.cuh, from an
`ncu --page source --print-source cuda,sass --csv` dump.  Regions are found from markers in the source file itself.
usage: ncu_regions_fx.py dump.csv path/to/slode_fixed.cuh"""
import collections, csv, re, sys

src = open(sys.argv[2]).read().split('\n')
def find(pat, start=0):
    for i in range(start, len(src)):
        if re.search(pat, src[i]):
            return i + 1
    return 10**9
marks = [
    (1, 'f2 / vector helpers'), (find(r'struct Pl \{'), 'pl.init (+sort)'), (find(r'void seek\('), 'pl.seek'),
    (find(r'void eval\(float te'), 'pl.eval'), (find(r'float warp_sum_scatter'), 'warp reductions'),
    (find(r'struct LatSmem'), 'small-net staging / chunks'), (find(r'fwd_smem_bytes'), 'tables / prologue'),
    (find(r'^fixed_fwd_kernel'), 'fwd kernel body'), (find(r'struct GradLayout'), 'layout'),
    (find(r'void events\('), 'sweep.events'), (find(r'void add\(float te'), 'sweep.add'),
    (find(r'bwd_smem_bytes'), 'bwd setup / prologue'), (find(r'for \(int i = T - 2; i >= 0'), 'bwd main loop (loads, stages, adjoint)'),
    (find(r'end of the sweep: per hidden unit'), 'finish: per-unit pass'), (find(r'auto outer = '), 'epilogue: small-net gradients'),
    (find(r'head biases: total'), 'bias sums / flush'), (find(r'struct Plan'), 'host'),
]
marks.sort()
def region(fn, ln):
    if not fn.startswith('slode_fixed'):
        return fn
    name = marks[0][1]
    for a, n in marks:
        if ln >= a:
            name = n
    return name
def f(x):
    try: return float(x.replace(',', ''))
    except Exception: return 0.0
rows = list(csv.reader(open(sys.argv[1])))
cur, hdr = '', None
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit(): continue
    a = agg[region(cur, int(r[0]))]
    a[0] += f(r[ci]); a[1] += f(r[si])
    for h, i in stall.items(): a[2][h] += f(r[i])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total warp instructions %.4g, samples %d' % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    top = ', '.join('%s %.0f%%' % (h[6:], 100 * c / max(v[1], 1)) for h, c in v[2].most_common(4))
    print('  %-42s %5.1f%% ins %5.1f%% smp | %s' % (k, 100 * v[0] / ti, 100 * v[1] / ts, top))
