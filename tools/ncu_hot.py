"""Top-sampled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output.
usage: python tools/ncu_hot.py src.csv <kernel index> [top N]"""
import csv, sys
lines = open(sys.argv[1]).read().split('\n')
idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
k = int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print(lines[idx[k]][:140])
rows = list(csv.reader(l for l in lines[idx[k] + 1: idx[k + 1]] if l.strip()))
hdr, rows = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[col['# Samples']]) for r in rows)
print("total samples", tot)
agg = {s: sum(int(r[col[s]]) for r in rows) for s in stall}
print({s: v for s, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][col['# Samples']]))[:top]
for i in sorted(order):
    r = rows[i]
    st = {s.replace('stall_', ''): int(r[col[s]]) for s in stall if int(r[col[s]])}
    print("%5d %6s %7s  %-70s %s" % (i, r[col['# Samples']], r[col['Instructions Executed']], r[col['Source']].strip()[:70], st))
