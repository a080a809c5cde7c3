"""Opcode histogram of a kernel (and of its innermost backward-branch loop) from the built .so: python tools/sass_loop.py fir_up"""
import collections
import re
import subprocess
import sys

out = subprocess.run(['cuobjdump', '-sass', 'next3d_b200/libnext3d_b200.so'], capture_output=True, text=True).stdout
cur, d = None, {}
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); d[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m and cur:
        d[cur].append((int(m.group(1), 16), re.sub(r'^@!?U?P\d\s+', '', m.group(2))))
for k, L in d.items():
    if sys.argv[1] in k:
        loops = []
        for a, t in L:
            m = re.search(r'BRA\S*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
            if m and int(m.group(1), 16) < a:
                loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
        print(k[-60:], 'total', len(L), 'loops', [(hex(b), hex(e), (e - b) // 16) for _, b, e in loops])
        for _, b, e in sorted(loops)[:3]:
            c = collections.Counter(t.split()[0] for a, t in L if b <= a <= e)
            print('  loop', hex(b), sum(c.values()), c.most_common(24))
