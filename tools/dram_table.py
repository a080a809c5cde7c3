"""ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch -> per-kernel DRAM traffic table
(JSON, read by bench.py for roofline.traffic).
Usage: python tools/dram_table.py gpurun_out/dram_r01.csv profiles/r01_dram_traffic.json"""
import collections
import csv
import json
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}


def main(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    h = rows[0]
    ki, mi, ui, vi, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value'), h.index('ID')
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[(int(r[ii]), r[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', ''))][r[mi]] = float(r[vi].replace(',', '')) * UNIT.get(r[ui], 1.0)
    keys = sorted(per.keys())
    starts = [i for i, (_, n) in enumerate(keys) if n == 'styles_kernel']        # one full generator step, like tools/launch_table.py
    if len(starts) >= 2:
        keys = keys[starts[0]:starts[1]]
    out = collections.OrderedDict()
    for (lid, name) in keys:
        m = per[(lid, name)]
        o = out.setdefault(name, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_us=0.0))
        o['launches'] += 1
        o['dram_read_bytes'] += m.get('dram__bytes_read.sum', 0.0)
        o['dram_write_bytes'] += m.get('dram__bytes_write.sum', 0.0)
        o['time_us'] += m.get('gpu__time_duration.sum', 0.0)
    for o in out.values():
        o['dram_bytes_per_launch'] = (o['dram_read_bytes'] + o['dram_write_bytes']) / o['launches']
    json.dump(dict(source=src, note='one generator forward (batch 8), ncu --clock-control none, cold L2 per launch', kernels=out), open(dst, 'w'), indent=1)
    for k, o in out.items():
        print(f"{k:32s} {o['launches']:4d} launches  {o['dram_read_bytes'] / 1e9:7.3f} GB read {o['dram_write_bytes'] / 1e9:7.3f} GB written  {o['time_us']:9.1f} us")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
