"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,...] --csv) into a per-kernel table for one generator step.
Usage: python tools/launch_table.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    h = next(r)
    ix = {n: i for i, n in enumerate(h)}
    unit = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}
    rows = [(re.sub(r'<unnamed>::|void ', '', row[ix['Kernel Name']]).split('(')[0], float(row[ix['Metric Value']].replace(',', '')) * unit.get(row[ix['Metric Unit']], 1e-3))
            for row in r if row[ix['Metric Name']] == 'gpu__time_duration.sum']      # the list may carry further metrics (DRAM bytes) per launch
    starts = [i for i, (n, _) in enumerate(rows) if n == 'styles_kernel']
    if len(starts) >= 2:
        rows = rows[starts[0]:starts[1]]
        print(f'one full step = launches between two styles_kernel launches: {len(rows)} launches')
    else:
        print(f'no complete step delimited by styles_kernel found; using all {len(rows)} rows')
    tot = sum(t for _, t in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print(f'sum of kernel durations (cold-cache, serialised under ncu): {tot / 1e3:.2f} ms')
    print(f'{"kernel":32s} {"launches":>8s} {"total us":>10s} {"share":>7s}')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{n:32s} {c:8d} {t:10.1f} {100 * t / tot:6.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
