"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total, average, share.
usage: launch_shares.py launches.csv [title]"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r'\(.*$', '', r[4]).replace('void ', '').strip()
    name = re.sub(r'^.*<unnamed>::', '', name) if name.startswith('<unnamed>') or '::<unnamed>::k_' in name else name
    a = agg.setdefault(name[:70], [0, 0.0])
    a[0] += 1; a[1] += float(r[14]) / 1e3
tot = sum(v[1] for v in agg.values())
print('# %s' % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print('\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print('| `%s` | %d | %.1f | %.1f | %.1f%% |' % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
