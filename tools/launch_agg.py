"""Aggregate an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('cgcn::', '')
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %4d %9.1f us %5.1f%%  avg %7.1f" % (k[:58], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
print("total %.1f us, %d launches" % (tot, sum(v[0] for v in agg.values())))
