"""Pack the reference's demo FLAME mesh + landmarks into a compact binary fixture.

Run once in the build container (the only place /root/reference exists):
    python tools/make_assets.py
Source data (inputs only, no code): data/demo/demo.obj ('v', 'vt', 'f a/b' lines, read the way
gen_samples_next3d.py:165-174 and pytorch3d.io.load_obj do) and data/demo/demo_kpt2d.txt
(np.loadtxt, gen_samples_next3d.py:176-178).
"""
import sys
import numpy as np

REF = '/root/reference/data/demo'


def parse_obj(path):
    v, vt, fv, ft = [], [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith('v '):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('vt '):
                vt.append([float(x) for x in line.split()[1:3]])
            elif line.startswith('f '):
                a, b = [], []
                for tok in line.split()[1:4]:
                    p = tok.split('/')
                    a.append(int(p[0]) - 1)
                    b.append(int(p[1]) - 1)
                fv.append(a)
                ft.append(b)
    return (np.asarray(v, np.float64), np.asarray(vt, np.float32),
            np.asarray(fv, np.int32), np.asarray(ft, np.int32))


def main():
    v, vt, fv, ft = parse_obj(f'{REF}/demo.obj')
    lms = np.loadtxt(f'{REF}/demo_kpt2d.txt')
    assert v.shape == (5023, 3) and vt.shape == (5118, 2) and fv.shape == (9976, 3) and lms.shape == (68, 3)
    out = 'next3d_b200/assets/flame_demo.npz'
    np.savez_compressed(out, verts=v.astype(np.float32), verts_uvs=vt, faces=fv, uvfaces=ft,
                        lms=lms.astype(np.float32))
    print('wrote', out)


if __name__ == '__main__':
    sys.exit(main())
