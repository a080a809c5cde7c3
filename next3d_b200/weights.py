"""Deterministic synthetic weights + FLAME topology buffers, keyed by the reference's state-dict names.

There are no public checkpoints in the reference tree (README.md:40 points at a download), so parity
tests and the benchmark use seeded random-init weights drawn from the same distributions the reference
constructors use (randn weights, networks_stylegan2.py:109,304; const, :524) -- except that biases and
noise strengths, which the reference initialises to exactly 0 (:308-309), are randomised so that bias /
noise bugs cannot hide.  Each tensor has its own generator seeded from
(seed, crc32(name)), so the dict is independent of construction order and can be loaded into the
reference's modules (`load_state_dict`) as well as into ours.
"""
import os
import zlib

import numpy as np
import torch

from . import config as _config

_ASSET = os.path.join(os.path.dirname(__file__), 'assets', 'flame_demo.npz')


def load_flame_demo():
    """-> dict(verts [5023,3] f32, verts_uvs [5118,2] f32, faces [9976,3] i32, uvfaces [9976,3] i32, lms [68,3] f32)."""
    with np.load(_ASSET) as z:
        return {k: z[k] for k in z.files}


def _dense_triangles(h, w, margin_x=2, margin_y=5):
    # generate_triangles: volumetric_rendering/renderer.py:475-491 (dead on the hot path, kept for the state dict)
    xs = np.arange(margin_x, w - 1 - margin_x)
    ys = np.arange(margin_y, h - 1 - margin_y)
    X, Y = np.meshgrid(xs, ys, indexing='ij')
    X, Y = X.reshape(-1), Y.reshape(-1)
    t0 = np.stack([Y * w + X, Y * w + X + 1, (Y + 1) * w + X], 1)
    t1 = np.stack([Y * w + X + 1, (Y + 1) * w + X + 1, (Y + 1) * w + X], 1)
    tri = np.stack([t0, t1], 1).reshape(-1, 3)
    return tri[:, [0, 2, 1]]


def topology_buffers(mesh=None):
    """Buffers registered at triplane_next3d.py:79-103."""
    mesh = mesh or load_flame_demo()
    uv = torch.from_numpy(mesh['verts_uvs'])[None]                       # raw_uvcoords [1,VT,2]
    uvfaces = torch.from_numpy(mesh['uvfaces'].astype(np.int64))[None]
    faces = torch.from_numpy(mesh['faces'].astype(np.int64))[None]
    uvc = torch.cat([uv, uv[:, :, 0:1] * 0. + 1.], -1)                   # :98
    uvc = uvc * 2 - 1
    uvc[..., 1] = -uvc[..., 1]                                           # :99
    face_uv = uvc[0][uvfaces[0]][None]                                   # face_vertices, renderer.py:444-463
    return {
        'dense_faces': torch.from_numpy(_dense_triangles(256, 256)).long()[None].contiguous(),
        'faces': faces, 'raw_uvcoords': uv, 'uvcoords': uvc, 'uvfaces': uvfaces, 'face_uvcoords': face_uv,
    }


def _gen(seed, name):
    g = torch.Generator(device='cpu')
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def make_state_dict(cfg, seed=0):
    """name -> CPU float32 tensor for every parameter/buffer of TriPlaneGenerator (config.param_spec)."""
    sd = {}
    topo = None
    filt = torch.tensor([1., 3., 3., 1.])
    filt = torch.outer(filt, filt)
    filt = filt / filt.sum()                                             # upfirdn2d.setup_filter, upfirdn2d.py:101-111
    for name, shape, kind, extra in _config.param_spec(cfg):
        if kind == 'topology':
            topo = topo or topology_buffers()
            sd[name] = topo[name]
            continue
        g = _gen(seed, name)
        if kind == 'weight':
            t = torch.randn(shape, generator=g) / float(extra.get('lr_mul', 1.0))
        elif kind == 'bias':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'affine_bias':
            t = 1.0 + torch.randn(shape, generator=g) * 0.1
        elif kind in ('noise_const', 'const'):
            t = torch.randn(shape, generator=g)
        elif kind == 'noise_strength':
            t = torch.rand(shape, generator=g) * 0.1
        elif kind == 'w_avg':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'filter':
            t = filt.clone()
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd


def demo_inputs(cfg, batch, seed=0, device='cpu', jitter=0.0):
    """Synthetic (z, c_cond, c_cam, v) following gen_samples_next3d.py:160-197:
    z = RandomState(seed+i).randn(512); cameras LookAtPoseSampler(pi/2 + yaw, pi/2 - 0.2, pivot [0,0,0.2], r 2.7)
    with yaw cycling (.4, 0, -.4); intrinsics focal 4.2647; v = demo verts ++ 68 landmarks."""
    from . import camera
    mesh = load_flame_demo()
    z = np.stack([np.random.RandomState(seed + i).randn(cfg.z_dim) for i in range(batch)]).astype(np.float32)
    yaws = [0.4, 0.0, -0.4]
    pivot = torch.tensor(cfg.rendering_kwargs.get('avg_camera_pivot', [0, 0, 0]), dtype=torch.float32)
    radius = cfg.rendering_kwargs.get('avg_camera_radius', 2.7)
    intr = camera.fov_to_intrinsics(18.837)
    cams, conds = [], []
    for i in range(batch):
        c2w = camera.look_at_pose(np.pi / 2 + yaws[i % 3], np.pi / 2 - 0.2, pivot, radius)
        cams.append(torch.cat([c2w.reshape(16), intr.reshape(9)]))
        c2w0 = camera.look_at_pose(np.pi / 2, np.pi / 2, pivot, radius)
        conds.append(torch.cat([c2w0.reshape(16), intr.reshape(9)]))
    v = torch.from_numpy(np.concatenate([mesh['verts'], mesh['lms']], 0))[None].repeat(batch, 1, 1)
    if jitter > 0:
        g = _gen(seed, 'vertex_jitter')
        v = v + torch.randn(v.shape, generator=g) * jitter
    return (torch.from_numpy(z).to(device), torch.stack(conds).to(device), torch.stack(cams).to(device),
            v.to(device))


def sampler_noise(cfg, batch, res=None, seed=0, depth=None, depth_importance=None):
    """Uniforms injected into both the oracle/reference and the kernel: u_coarse [N,M,Dc,1], u_fine [N*M,Df]."""
    res = res or cfg.neural_rendering_resolution
    dc = depth or cfg.rendering_kwargs['depth_resolution']
    df = depth_importance or cfg.rendering_kwargs['depth_resolution_importance']
    g = _gen(seed, 'sampler_noise')
    return (torch.rand(batch, res * res, dc, 1, generator=g), torch.rand(batch * res * res, df, generator=g))
