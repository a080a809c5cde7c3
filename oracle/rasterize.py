"""ctypes wrapper around oracle_c.c (CPU oracle: rasterizer + flood fill).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE, 'liboracle_c.so'])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, 'liboracle_c.so')
        if not os.path.exists(so):
            build()
        _lib = ctypes.CDLL(so)
        _lib.n3d_oracle_rasterize.restype = None
        _lib.n3d_oracle_floodfill.restype = None
    return _lib


def rasterize_meshes(verts, faces, H, W):
    """verts [N,V,3] f32 (pytorch3d NDC), faces [N,F,3] int -> pix_to_face [N,H,W] i64, zbuf [N,H,W], bary [N,H,W,3]."""
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    N, V, _ = verts.shape
    F = faces.shape[1]
    p2f = np.empty((N, H, W), np.int64)
    zbuf = np.empty((N, H, W), np.float32)
    bary = np.empty((N, H, W, 3), np.float32)
    P = ctypes.c_void_p
    lib().n3d_oracle_rasterize(P(verts.ctypes.data), P(faces.ctypes.data), N, V, F, H, W,
                               P(p2f.ctypes.data), P(zbuf.ctypes.data), P(bary.ctypes.data))
    return p2f, zbuf, bary


def floodfill_(img, lo=0.0, up=254.0, newval=255.0):
    """In-place cv2.floodFill(seed=(0,0), FLOODFILL_FIXED_RANGE, 4-connected) on a [H,W] f32 array."""
    assert img.dtype == np.float32 and img.flags['C_CONTIGUOUS'] and img.ndim == 2
    lib().n3d_oracle_floodfill(ctypes.c_void_p(img.ctypes.data), img.shape[0], img.shape[1],
                               ctypes.c_float(lo), ctypes.c_float(up), ctypes.c_float(newval))
    return img
