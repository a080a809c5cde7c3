"""Join an ncu source page (SASS rows with samples / executed counts) with `nvdisasm --print-line-info-inline` output to attribute
executed warp instructions and stall samples to CUDA source lines of one file (outermost inlining site in that file).
Usage: python tools/sass_lines.py <ncu source csv> <nvdisasm listing> <kernel substring> <file.cu> [top] [range,range,...]
  ranges like 160-205:gather  attribute line ranges to named regions."""
import csv, re, sys
from collections import defaultdict


def parse_sass(sass, kernel, fname):
    lines = open(sass).read().split('\n')
    infn = False; group = []; per_inst = []; last = None
    for ln in lines:
        if ln.startswith('.text.') or ln.lstrip().startswith('.section'):
            infn = kernel in ln and '.text.' in ln
            continue
        if not infn:
            continue
        if '//## File' in ln:
            group.append(ln)
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            if group:
                cand = None
                for g in group:
                    for m in re.finditer(r'"([^"]+)", line (\d+)', g):
                        if m.group(1).endswith(fname):
                            cand = int(m.group(2))
                last = cand if cand is not None else last
                group = []
            per_inst.append(last)
    return per_inst


def main(src_csv, sass, kernel, fname, top=40, ranges=''):
    rows = list(csv.reader(open(src_csv)))
    h = rows[1]; d = rows[2:]
    iex, isamp = h.index('Instructions Executed'), h.index('# Samples')
    per_inst = parse_sass(sass, kernel, fname)
    print('sass instructions', len(per_inst), 'ncu rows', len(d))
    n = min(len(per_inst), len(d))
    ex = defaultdict(int); sm = defaultdict(int)
    for i in range(n):
        ex[per_inst[i]] += int(d[i][iex]); sm[per_inst[i]] += int(d[i][isamp])
    tot = sum(ex.values()); tots = sum(sm.values())
    print('total executed', tot, 'samples', tots)
    if ranges:
        for spec in ranges.split(','):
            rg, name = spec.split(':'); a, b = map(int, rg.split('-'))
            e = sum(v for k, v in ex.items() if k is not None and a <= k <= b); s = sum(v for k, v in sm.items() if k is not None and a <= k <= b)
            print(f'  {name:24s} lines {a:4d}-{b:4d}  exec {e:>12d} {100*e/tot:5.1f}%  samples {s:>7d} {100*s/tots:5.1f}%')
    for k, v in sorted(ex.items(), key=lambda x: -x[1])[:top]:
        print(f'line {str(k):6s} exec {v:>12d} {100*v/tot:5.1f}%   samples {sm[k]:>7d} {100*sm[k]/tots:5.1f}%')


if __name__ == '__main__':
    a = sys.argv
    main(a[1], a[2], a[3], a[4], int(a[5]) if len(a) > 5 else 40, a[6] if len(a) > 6 else '')
