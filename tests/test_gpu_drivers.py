"""GPU tests of the callers either side of the hot path: the batched frame driver against a
per-frame loop with identical injected sampler noise, and the shape-extraction grid (in-kernel voxel coordinates, fused flip +
trim) against the oracle's run_model and against sample_mixed on the reference's full coordinate tensor."""
import math

import numpy as np
import pytest
import torch

from tests.helpers import range_rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def K():
    from next3d_b200 import kernels
    return kernels


@pytest.fixture(scope='module')
def tinyG():
    from next3d_b200 import config, weights
    from next3d_b200.triplane_next3d import TriPlaneGenerator
    cfg = config.tiny_config(512)
    return cfg, TriPlaneGenerator.from_config(cfg, weights.make_state_dict(cfg, 3), device=DEV)


def test_render_frames_equals_per_frame_loop(tinyG):
    """drivers.render_frames (batches of 2, last batch padded) == one synthesis call per frame, pixel for pixel, when every
    frame carries its own sampler noise (gen_videos_next3d.py:128-171 renders one frame per call)."""
    from next3d_b200 import drivers, weights
    cfg, G = tinyG
    z, cc, c, v = weights.demo_inputs(cfg, 2, seed=3)
    ws = G.mapping(z.to(DEV), cc.to(DEV), truncation_psi=0.7, truncation_cutoff=14)
    F = 5
    wsf = drivers.interpolate_ws(ws.cpu(), 3, wraps=1)[:F]
    cams = drivers.orbit_camera_params(F, torch.tensor([0, 0, 0.2]), 2.7)
    R = G.neural_rendering_resolution
    D, Df = G.rendering_kwargs['depth_resolution'], G.rendering_kwargs['depth_resolution_importance']
    g = torch.Generator().manual_seed(11)
    u_c = torch.rand(F, R * R, D, 1, generator=g).to(DEV)
    u_f = torch.rand(F, R * R, Df, generator=g).to(DEV)
    got = list(drivers.render_frames(G, wsf, cams, v[:1], batch=2, sampler_noise=(u_c, u_f)))
    assert len(got) == F and got[0].shape == (cfg.img_resolution, cfg.img_resolution, 3) and got[0].dtype == np.uint8
    for f in range(F):
        ref = G.synthesis(wsf[f:f + 1].to(DEV).float(), cams[f:f + 1].to(DEV), v[:1].to(DEV), noise_mode='const',
                          sampler_noise=(u_c[f:f + 1], u_f[f]))['image']
        ref = drivers.to_uint8_hwc(ref)[0].cpu().numpy()
        assert np.array_equal(got[f], ref), f'frame {f}: max |diff| {np.abs(got[f].astype(int) - ref.astype(int)).max()}'


def test_render_frames_graph_path(tinyG):
    """The CUDA-graph replay path of the driver (in-kernel RNG): right count, shape and dtype, frames differ along the orbit."""
    from next3d_b200 import drivers, weights
    cfg, G = tinyG
    z, cc, c, v = weights.demo_inputs(cfg, 1, seed=3)
    ws = G.mapping(z.to(DEV), cc.to(DEV), truncation_psi=0.7, truncation_cutoff=14)
    F = 5
    wsf = ws.cpu().expand(F, -1, -1)
    cams = drivers.orbit_camera_params(F, torch.tensor([0, 0, 0.2]), 2.7)
    G.use_cuda_graph = True
    try:
        got = list(drivers.render_frames(G, wsf, cams, v[:1], batch=2, seed=7))
    finally:
        G.use_cuda_graph = False
    assert len(got) == F and all(f.shape == (cfg.img_resolution, cfg.img_resolution, 3) and f.dtype == np.uint8 for f in got)
    assert np.abs(got[0].astype(int) - got[2].astype(int)).max() > 0


@pytest.mark.parametrize('R,head,count', [(64, 0, None), (40, 12345, 30000)])
def test_sample_grid_vs_oracle(K, R, head, count):
    """n3d_sample_grid (voxel centres generated in the kernel, flip + trim fused, border voxels skipped) == oracle run_model on the
    reference's create_samples coordinates followed by the script's flip + trim (gen_samples_next3d.py:80-102, 208-238)."""
    from next3d_b200 import drivers
    from oracle import renderer as orr
    g = torch.Generator().manual_seed(5)
    planes = torch.randn(1, 3, 32, 64, 64, generator=g)
    sd = {'decoder.net.0.weight': torch.randn(64, 32, generator=g), 'decoder.net.0.bias': torch.randn(64, generator=g) * 0.1,
          'decoder.net.2.weight': torch.randn(33, 64, generator=g), 'decoder.net.2.bias': torch.randn(33, generator=g) * 0.1}
    dec = ((sd['decoder.net.0.weight'] / math.sqrt(32)).to(DEV).contiguous(), sd['decoder.net.0.bias'].to(DEV),
           (sd['decoder.net.2.weight'] / math.sqrt(64)).to(DEV).contiguous(), sd['decoder.net.2.bias'].to(DEV))
    pad = int(30 * R / 256)
    coords = drivers.create_samples(R, 1.0)                                   # the reference's full [1, R^3, 3] tensor (CPU)
    _, sig_ref = orr.run_model(sd, planes, coords, {'box_warp': 1})
    ref = drivers.trim_sigma_grid(sig_ref.reshape(R, R, R).clone(), R)
    out = torch.full((R, R, R), float('nan'), device=DEV)
    cnt = R ** 3 - head if count is None else count
    K.sample_grid(planes.permute(0, 1, 3, 4, 2).contiguous().to(DEV)[0], R, 1.0, 1.0, dec, out, head=head, count=cnt, pad=pad)
    out = out.cpu()
    # only the flat indices head .. head+count-1 are written; map them through the flip to a mask
    idx = torch.arange(head, head + cnt)
    a, b, c = idx // (R * R), (idx // R) % R, idx % R
    mask = torch.zeros(R, R, R, dtype=torch.bool)
    mask[R - 1 - a, b, c] = True
    assert torch.isnan(out[~mask]).all() and not torch.isnan(out[mask]).any()
    border = ref == -1000.0
    assert torch.equal(out[mask & border], ref[mask & border])
    assert range_rel_err(out[mask & ~border], ref[mask & ~border]) < 3e-5          # bf16x3 tensor-core decoder (~2^-16 per product)


def test_extract_sigma_grid_equals_sample_mixed(tinyG):
    """Row f2 end to end: planes computed once + grid kernel == sample_mixed on the reference's coordinate tensor + script trim."""
    from next3d_b200 import drivers, weights
    cfg, G = tinyG
    z, cc, c, v = weights.demo_inputs(cfg, 1, seed=3)
    ws = G.mapping(z.to(DEV), cc.to(DEV), truncation_psi=0.7, truncation_cutoff=14)
    R = 32
    grid = drivers.extract_sigma_grid(G, ws, v.to(DEV), shape_res=R)
    coords = drivers.create_samples(R, G.rendering_kwargs['box_warp']).to(DEV)
    ref = G.sample_mixed(coords, None, ws, v.to(DEV), noise_mode='const')['sigma'].reshape(R, R, R)
    ref = drivers.trim_sigma_grid(ref.clone(), R).cpu().numpy()
    assert grid.shape == (R, R, R) and grid.dtype == np.float32
    assert np.array_equal(grid, ref)


def test_reenact_frames_and_pack(tinyG, tmp_path):
    """Row f1/f3: reenactment driver over a synthetic driving directory (dataset.json labels, per-frame .obj + landmarks parsed by
    the native parsers on worker threads) == per-frame synthesis with the script's camera smoothing; the binary pack gives the
    same frames without parsing."""
    import json
    from next3d_b200 import drivers, inputs, weights
    cfg, G = tinyG
    z, cc, c, v = weights.demo_inputs(cfg, 1, seed=3)
    ws = G.mapping(z.to(DEV), cc.to(DEV), truncation_psi=0.7, truncation_cutoff=14)
    n = 6
    rng = np.random.RandomState(1)
    labels = []
    for k in range(n):
        name = f'{k:04d}'
        (tmp_path / f'{name}.png').write_bytes(b'')
        vk = v[0].numpy() + rng.randn(*v[0].shape).astype(np.float32) * 1e-4
        (tmp_path / f'{name}.obj').write_text(''.join(f'v {float(a)!r} {float(b)!r} {float(c_)!r}\n' for a, b, c_ in vk[:5023]))
        (tmp_path / f'{name}_kpt2d.txt').write_text('\n'.join(f'{float(a)!r} {float(b)!r} {float(c_)!r}' for a, b, c_ in vk[5023:]) + '\n')
        cam = c[0].numpy().astype(np.float64)
        cam[3] += 0.01 * k
        labels.append([f'{name}.png', cam.tolist()])
    (tmp_path / 'dataset.json').write_text(json.dumps({'labels': labels}))
    root = str(tmp_path)
    sch = drivers.reenact_schedule(root, 100)
    F = len(sch['ks'])
    assert F == n - 2
    R = G.neural_rendering_resolution
    D, Df = G.rendering_kwargs['depth_resolution'], G.rendering_kwargs['depth_resolution_importance']
    g = torch.Generator().manual_seed(3)
    u_c, u_f = torch.rand(F, R * R, D, 1, generator=g).to(DEV), torch.rand(F, R * R, Df, generator=g).to(DEV)
    got = list(drivers.reenact_frames(G, ws, root, 100, batch=3, sampler_noise=(u_c, u_f)))
    assert len(got) == F
    verts = []
    for f in range(F):
        vf = inputs.load_frame(sch['obj_paths'][f], sch['lms_paths'][f])
        verts.append(vf[0].numpy())
        ref = G.synthesis(ws, sch['cams'][f:f + 1].to(DEV), vf.to(DEV), noise_mode='const', sampler_noise=(u_c[f:f + 1], u_f[f]))['image']
        assert np.array_equal(got[f], drivers.to_uint8_hwc(ref)[0].cpu().numpy()), f
    pack = str(tmp_path / 'clip.n3dpack')
    inputs.write_frame_pack(pack, np.stack(verts), sch['cams'].numpy(), ids=sch['ids'])
    got2 = list(drivers.reenact_frames(G, ws, pack, 100, batch=3, sampler_noise=(u_c, u_f)))
    assert len(got2) == F and all(np.array_equal(a, b) for a, b in zip(got, got2))


def test_interpolate_ws_on_device(tinyG):
    from next3d_b200 import drivers
    g = torch.Generator().manual_seed(2)
    ws_key = torch.randn(3, 28, 512, generator=g)
    ref = drivers.interpolate_ws(ws_key, 7, wraps=2)
    got = drivers.interpolate_ws_device(ws_key.to(DEV), 7, wraps=2)
    assert got.shape == ref.shape and range_rel_err(got.cpu(), ref) < 2e-6
