"""Per-frame inputs of the inference scripts: the FLAME mesh (.obj) and the 68 2-D landmarks
(*_kpt2d.txt) that `gen_samples_next3d.py:165-178`, `gen_videos_next3d.py:118-131` and `reenact_avatar_next3d.py:128-141`
parse in Python for every frame, plus a prefetcher that keeps parsed frames in pinned host memory ahead of the generator.

The parsers are native (`n3d_parse_obj_vertices`, `n3d_parse_float_table` in libnext3d_b200.so): one pass over the file, values
bit-identical to the reference's `float(token)` -> float64 -> `.float()`.
"""
import concurrent.futures
import ctypes as C

import numpy as np
import torch

from . import _lib


def _read(path_or_bytes):
    return bytes(path_or_bytes) if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, 'rb').read()


def _err():
    return ValueError(_lib.lib.n3d_last_error().decode())


def load_obj_vertices(path_or_bytes):
    """-> float32 [V, 3]: the `v ` lines of a Wavefront .obj (gen_samples_next3d.py:165-174)."""
    data = _read(path_or_bytes)
    cap = len(data) // 6 + 1                        # a vertex line is at least "v 0 0 0": one pass, no sizing pass
    out = np.empty((cap, 3), np.float32)
    n = C.c_int64(0)
    if _lib.lib.n3d_parse_obj_vertices(data, len(data), out.ctypes.data, cap, C.byref(n)) != 0:
        raise _err()
    return out[: n.value].copy() if n.value * 4 < cap else out[: n.value]


def load_float_table(path_or_bytes):
    """-> float32 [rows, cols] (1-D for a single row, like np.loadtxt): whitespace-separated numbers, '#' comments."""
    data = _read(path_or_bytes)
    cap = len(data) // 2 + 1                        # a number and its separator take at least two bytes
    out = np.empty(cap, np.float32)
    n, cols = C.c_int64(0), C.c_int64(0)
    if _lib.lib.n3d_parse_float_table(data, len(data), out.ctypes.data, cap, C.byref(n), C.byref(cols)) != 0:
        raise _err()
    out = out[: n.value].copy()
    if cols.value == 0:
        return out
    rows = n.value // cols.value
    return out.reshape(rows, cols.value) if rows > 1 else out


def load_frame(obj_path, lms_path=None, pin=False):
    """-> `v` as `TriPlaneGenerator.synthesis` takes it: float32 [1, V (+ 68), 3] = vertices, then the landmarks
    (gen_samples_next3d.py:165-178: `torch.cat((v, lms), 1)`).  pin=True returns page-locked memory (async H2D)."""
    parts = [load_obj_vertices(obj_path)]
    if lms_path is not None:
        lms = load_float_table(lms_path)
        parts.append(np.asarray(lms, np.float32).reshape(-1, parts[0].shape[1]))
    t = torch.from_numpy(np.concatenate(parts, 0))[None]
    return t.pin_memory() if pin else t


class FramePrefetcher:
    """Iterate over (obj_path, lms_path) pairs, parsing `depth` frames ahead on worker threads (the parsers release the GIL
    inside the native call) and yielding pinned `[1, V+68, 3]` tensors in order."""

    def __init__(self, frames, depth=4, workers=2, pin=None):
        self.frames = list(frames)
        self.depth = max(1, depth)
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=workers)

    def __len__(self):
        return len(self.frames)

    def __iter__(self):
        pending = []
        it = iter(self.frames)
        try:
            for _ in range(self.depth):
                f = next(it, None)
                if f is None:
                    break
                pending.append(self.pool.submit(load_frame, f[0], f[1], self.pin))
            while pending:
                out = pending.pop(0).result()
                f = next(it, None)
                if f is not None:
                    pending.append(self.pool.submit(load_frame, f[0], f[1], self.pin))
                yield out
        finally:
            for p in pending:
                p.cancel()
