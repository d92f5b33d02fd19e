"""Per-kernel times of ONE partitioned PCG iteration from an ncu launch list (csv) of profiles/r02_dist_iteration.py:
the launches between the xr_update kernels of two consecutive iterations, summed per kernel name and listed in order.
    python profiles/parse_dist_launches.py launches.csv parts"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
seq = [(r[ki].split('(')[0].replace('mwx::<unnamed>::', '').replace('void ', ''), float(r[vi].replace(',', '')) / 1e3, r[gi]) for r in rows[1:]]
idx = [i for i, (k, _, _) in enumerate(seq) if k.startswith('k_xr_update')]
a, b = idx[-1 - 2 * parts] , idx[-1 - parts]
tot = 0; by = collections.OrderedDict()
for k, t, g in seq[a + 1:b + 1]:
    print(f"{k[:60]:60s} {t:9.1f} us grid {g}"); tot += t
    e = by.setdefault(k[:60], [0, 0.0]); e[0] += 1; e[1] += t
print('iteration total us', tot, 'launches', b - a)
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} x{c:4d} {t:10.1f} us  {100 * t / tot:5.1f} %")
