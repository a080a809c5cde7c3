"""Frames/s of the per-frame input parsing: the reference's Python loop + np.loadtxt vs the native parsers.
Usage: python tools/bench_parse.py [obj] [kpt]   (defaults: the reference's demo frame if present, else a synthetic 5023-vertex mesh)"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from next3d_b200 import inputs  # noqa: E402


def ref_frame(obj, kpt):
    """The scripts' per-frame parsing (gen_samples_next3d.py:165-178) restated in Python: float() of every token of the 'v ' lines,
    np.loadtxt for the landmarks, torch.cat."""
    rows = [[float(tok) for tok in ln.split()[1:]] for ln in open(obj, 'r').read().split('\n') if ln[:2] == 'v ']
    v = torch.from_numpy(np.array(rows).reshape(-1, 3)).float()[None]
    return torch.cat((v, torch.from_numpy(np.loadtxt(kpt)).float()[None]), 1)


def main():
    if len(sys.argv) >= 3:
        obj, kpt = sys.argv[1:3]
    elif os.path.exists('/root/reference/data/demo/demo.obj'):
        obj, kpt = '/root/reference/data/demo/demo.obj', '/root/reference/data/demo/demo_kpt2d.txt'
    else:
        d = tempfile.mkdtemp()
        obj, kpt = os.path.join(d, 'm.obj'), os.path.join(d, 'k.txt')
        rng = np.random.default_rng(0)
        with open(obj, 'w') as f:
            for x in rng.standard_normal((5023, 3)):
                f.write('v %.6f %.6f %.6f\n' % tuple(x))
            for _ in range(9976):
                f.write('f 1/1 2/2 3/3\n')
        np.savetxt(kpt, rng.standard_normal((68, 3)))
    assert torch.equal(ref_frame(obj, kpt), inputs.load_frame(obj, kpt))
    for name, fn, reps in (('reference python loop + np.loadtxt', lambda: ref_frame(obj, kpt), 20), ('native parsers (inputs.load_frame)', lambda: inputs.load_frame(obj, kpt), 200)):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        print(f'{name:40s} {dt * 1e3:8.3f} ms / frame  {1 / dt:9.1f} frames/s')
    frames = [(obj, kpt)] * 256
    t0 = time.perf_counter()
    n = sum(1 for _ in inputs.FramePrefetcher(frames, depth=8, workers=4, pin=False))
    dt = (time.perf_counter() - t0) / n
    print(f'{"FramePrefetcher, 4 worker threads":40s} {dt * 1e3:8.3f} ms / frame  {1 / dt:9.1f} frames/s')


if __name__ == '__main__':
    main()
