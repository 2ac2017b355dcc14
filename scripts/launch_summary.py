#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_summary.py launches.csv [out.md]"""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r['Kernel Name'].split('(')[0].replace('void ', '')
        name = name.split('<unnamed>::')[-1] if '<unnamed>::' in name else name
        agg[name[-70:]][0] += 1
        agg[name[-70:]][1] += float(r['Metric Value'])
    tot = sum(v[1] for v in agg.values())
    out = [f'{len(rows)} launches, {tot / 1e3:.1f} us of device time (cold-cache, serialised: compare shares)', '',
           '| kernel | launches | total us | share |', '|---|---|---|---|']
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.002:
            continue
        out.append(f'| `{n}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |')
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
