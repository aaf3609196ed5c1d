"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
        name = re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name']))
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
        tot += v
    print(f"# {path}: {sum(c for c, _ in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / tot * 100:5.1f}% {t:10.1f}us n={c:4d} avg={t / c:8.1f}us  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
