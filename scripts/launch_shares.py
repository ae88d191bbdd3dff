"""Splits the ncu launch list of scripts/ncu_target.py --eager into its phases (a 1-element add kernel marks each boundary) and
writes per-phase kernel shares: kernel, launches, total_us, share.
    python scripts/launch_shares.py gpurun_out/r02b_launches_raw.csv profiles/r02b_launches"""
import csv
import sys
from collections import OrderedDict

raw, prefix = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(raw)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
kn, mv, mn = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
launches = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hdr + 1:] if r[mn] == 'gpu__time_duration.sum']
unit = rows[hdr + 1][h.index('Metric Unit')]
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1e-3)
marks = [i for i, (k, _) in enumerate(launches) if 'vectorized_elementwise_kernel' in k and 'CUDAFunctorOnSelf_add' in k]   # marker.add_(1.0) on a 1-element tensor
names = ['uncached_step', 'cache_fill', 'cached_step']
marks = marks[-3:]
for j, name in enumerate(names):
    seg = launches[marks[j] + 1:(marks[j + 1] if j + 1 < 3 else len(launches))]
    agg = OrderedDict()
    for k, v in seg:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    with open('%s_%s.csv' % (prefix, name), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_us', 'share'])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(t, 2), round(t / tot, 4)])
    print(name, 'kernel time %.0f us in %d launches' % (tot, len(seg)))
