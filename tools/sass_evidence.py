"""Count Blackwell-specific SASS mnemonics per kernel in the built library -> profiles/rNN_sass_evidence.txt
Usage: python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import re
import subprocess

PAT = re.compile(r'^(UTCHMMA|UTMALDG|UBLKCP|LDTM|UTCBAR|UTCATOMSWS|SYNCS|STG\.E\.ENL2\.256|LDG\.E\.ENL2\.256|FFMA2|FMUL2|FADD2|F2FP\.BF16)')
out = subprocess.run(['cuobjdump', '-sass', 'next3d_b200/libnext3d_b200.so'], capture_output=True, text=True).stdout
cur, cnt = None, collections.Counter()
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = re.sub(r'^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_', '', m.group(1))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m and cur and PAT.match(m.group(1)):
        cnt[(cur, m.group(1))] += 1
print('# cuobjdump -sass next3d_b200/libnext3d_b200.so : count of Blackwell-specific SASS mnemonics per kernel')
print('# (UTCHMMA = tcgen05.mma, UTMALDG = TMA cp.async.bulk.tensor, UBLKCP = 1-D cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,')
print('#  SYNCS = mbarrier, STG/LDG.E.ENL2.256 = 256-bit global accesses, FFMA2/FMUL2/FADD2 = packed fp32x2, F2FP.BF16 = packed bf16 convert)')
for (k, op), n in sorted(cnt.items()):
    print(f'{n:5d} {k[:70]:70s} {op}')
