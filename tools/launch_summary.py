"""per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):  python tools/launch_summary.py LIST.csv "title" > OUT.txt"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r'^void ', '', r[4]).split('(')[0]
    unit, val = r[-2], float(r[-1])
    ms = val / 1e6 if unit in ('ns', 'nsecond') else val / 1e3 if unit in ('us', 'usecond') else val
    tot[name][0] += 1; tot[name][1] += ms
allms = sum(v[1] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print('%d launches, %.1f ms of kernel time (ncu per-launch times: cold cache, serialised)\n' % (len(rows), allms))
print('%-100s %7s %12s %7s' % ('kernel', 'count', 'ms', 'share'))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-100s %7d %12.3f %6.2f%%' % (k[:100], v[0], v[1], 100 * v[1] / allms))
