"""Row f1 of SURVEY.md section 8 (batched frame drivers): the frame schedule and the batching logic against the reference's own
per-frame code (gen_videos_next3d.py:96-171, camera_utils.py:68-86).  CPU only; the generator is a stub here, the real one is
covered by tests/test_gpu_generator.py (batch invariance)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from next3d_b200 import camera, drivers

REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'camera_utils.py')), reason='reference tree not present (GPU box)')


@needs_ref
def test_cameras_match_reference_sampler():
    sys.path.insert(0, REF)
    try:
        import camera_utils
    finally:
        sys.path.remove(REF)
    lookat = torch.tensor([0, 0, 0.2])
    for h, v in [(math.pi / 2, math.pi / 2), (math.pi / 2 + 0.4, math.pi / 2 - 0.2), (0.3, 1e-7), (3.0, 3.2)]:
        ref = camera_utils.LookAtPoseSampler.sample(h, v, lookat, radius=2.7)[0]
        assert torch.equal(camera.look_at_pose(h, v, lookat, 2.7), ref)
    assert torch.equal(camera.fov_to_intrinsics(18.837), camera_utils.FOV_to_intrinsics(18.837))
    # the orbit of gen_videos_next3d.py:128-140, frame by frame
    F = 24
    c = drivers.orbit_camera_params(F, lookat, 2.7)
    intr = torch.tensor([[4.2647, 0, 0.5], [0, 4.2647, 0.5], [0, 0, 1]])
    for f in range(F):
        pose = camera_utils.LookAtPoseSampler.sample(3.14 / 2 + 0.35 * np.sin(2 * 3.14 * f / (F // 2)),
                                                      3.14 / 2 - 0.05 + 0.25 * np.cos(2 * 3.14 * f / (F // 2)), lookat, radius=2.7)
        assert torch.equal(c[f], torch.cat([pose.reshape(-1, 16), intr.reshape(-1, 9)], 1)[0])


def test_interpolated_latents_match_per_frame_calls():
    import scipy.interpolate
    g = torch.Generator().manual_seed(0)
    K, w_frames, wraps = 3, 5, 2
    ws = torch.randn(K, 28, 16, generator=g)
    got = drivers.interpolate_ws(ws, w_frames, wraps)
    x = np.arange(-K * wraps, K * (wraps + 1))                            # gen_videos_next3d.py:112-114
    interp = scipy.interpolate.interp1d(x, np.tile(ws.numpy(), [wraps * 2 + 1, 1, 1]), kind='cubic', axis=0)
    for f in range(K * w_frames):
        assert np.array_equal(got[f].numpy(), interp(f / w_frames))      # :143-144
    assert np.allclose(got[::w_frames].numpy(), ws.numpy(), atol=1e-6)    # passes through the key frames


@needs_ref
def test_uint8_conversion_is_layout_grid():
    layout_grid = _reference_scripts()[1].layout_grid                      # gen_videos_next3d.py:35-49
    g = torch.Generator().manual_seed(1)
    img = torch.randn(2, 3, 8, 8, generator=g) * 1.5
    got = drivers.to_uint8_hwc(img)
    for k in range(2):
        assert np.array_equal(got[k].numpy(), layout_grid(img[k:k + 1], grid_w=1, grid_h=1))


class _StubG(torch.nn.Module):
    """Deterministic stand-in with the generator's call signature: pixel = f(ws, c, v) per sample, independent of the batch."""
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def synthesis(self, ws, c, v, noise_mode='const', seed=None):
        self.calls.append(ws.shape[0])
        s = ws.mean(dim=(1, 2)) + c.mean(dim=1) + v.mean(dim=(1, 2))
        img = torch.tanh(s)[:, None, None, None] * torch.linspace(-1, 1, 48).reshape(1, 3, 4, 4)
        return {'image': img, 'image_depth': s[:, None, None, None].expand(-1, 1, 4, 4) + torch.arange(16.).reshape(1, 1, 4, 4)}


@pytest.mark.parametrize('F,batch,mesh', [(10, 4, 'static'), (8, 8, 'per_frame'), (5, 3, 'iter'), (1, 8, 'static')])
def test_batched_frames_equal_per_frame_loop(F, batch, mesh):
    g = torch.Generator().manual_seed(2)
    ws, cams = torch.randn(F, 28, 16, generator=g), torch.randn(F, 25, generator=g)
    vf = torch.randn(F, 7, 3, generator=g)
    G = _StubG()
    verts = {'static': vf[:1], 'per_frame': vf, 'iter': (vf[i:i + 1] for i in range(F))}[mesh]
    got = list(drivers.render_frames(G, ws, cams, verts, batch=batch, device='cpu'))
    assert len(got) == F and all(n == batch for n in G.calls) and len(G.calls) == -(-F // batch)
    for f in range(F):
        v = vf[:1] if mesh == 'static' else vf[f:f + 1]
        ref = drivers.to_uint8_hwc(_StubG().synthesis(ws[f:f + 1], cams[f:f + 1], v)['image'])[0].numpy()
        assert np.array_equal(got[f], ref)
    depth = list(drivers.render_frames(_StubG(), ws, cams, vf[:1], batch=batch, image_mode='image_depth', device='cpu'))
    assert len(depth) == F and depth[0].shape == (4, 4, 3) and depth[0].max() == 255


# ------------------------------------------------------------------------------------------------ row f2: shape extraction
def _reference_scripts():
    """The reference's own gen_samples_next3d / gen_videos_next3d modules (through the oracle's import shims)."""
    from oracle import ref_shim
    ref_shim.import_reference()
    with ref_shim.in_scratch():
        import gen_samples_next3d
        import gen_videos_next3d
    return gen_samples_next3d, gen_videos_next3d


@needs_ref
@pytest.mark.parametrize('N', [17, 64])
def test_chunked_samples_are_slices_of_the_reference_grid(N):
    ref = _reference_scripts()[0].create_samples(N=N, voxel_origin=[0, 0, 0], cube_length=1.0)[0]
    assert torch.equal(drivers.create_samples(N, 1.0), ref)
    for head, n in [(0, 1000), (N ** 3 - 777, 777), (12345 % N ** 3, 2048)]:
        n = min(n, N ** 3 - head)
        assert torch.equal(drivers.create_samples(N, 1.0, head, n), ref[:, head:head + n])


def test_large_index_rounding_like_the_reference():
    """Indices above 2^24 are rounded by .float() in the reference; the chunked generator must reproduce that, not fix it."""
    N, head, n = 512, 2 ** 26 + 12345, 4096
    idx = torch.arange(head, head + n, dtype=torch.int64)
    y = (idx.float() / N) % N
    got = drivers.create_samples(N, 1.0, head, n)[0]
    assert torch.equal(got[:, 1], y * (1.0 / (N - 1)) + (-0.5))
    assert (idx.float().long() != idx).any()                              # the rounding really happens in this range


def test_trim_matches_script():
    """flip along axis 0, then a border of int(30 * R / 256) voxels on all six faces set to -1000 (gen_samples_next3d.py:226-238)."""
    R = 64
    g = torch.Generator().manual_seed(3)
    sig = torch.randn(R, R, R, generator=g)
    ref = sig.numpy()[::-1].copy()
    pad = int(30 * R / 256)
    inner = np.zeros((R, R, R), bool)
    inner[pad:R - pad, pad:R - pad, pad:R - pad] = True
    ref[~inner] = -1000
    assert np.array_equal(drivers.trim_sigma_grid(sig.clone(), R).numpy(), ref)


def _make_drive_root(tmp_path, n=6, with_lms=True):
    import json
    rng = np.random.RandomState(0)
    labels = []
    for k in range(n):
        name = f'{k:04d}'
        (tmp_path / f'{name}.png').write_bytes(b'')
        v = rng.randn(11, 3)
        (tmp_path / f'{name}.obj').write_text(''.join(f'v {a:.6f} {b:.6f} {c:.6f}\n' for a, b, c in v) + 'f 1 2 3\n')
        if with_lms:
            (tmp_path / f'{name}_kpt2d.txt').write_text('\n'.join(f'{a:.5f} {b:.5f} {c:.5f}' for a, b, c in rng.randn(68, 3)) + '\n')
        labels.append([f'{name}.png', rng.randn(25).tolist()])
    (tmp_path / 'dataset.json').write_text(json.dumps({'labels': labels}))
    return labels


def test_reenact_schedule_matches_the_script(tmp_path):
    """reenact_avatar_next3d.py:125-160 restated with its own expressions: frame selection, file names, smoothed cameras."""
    import glob
    import os
    labels = _make_drive_root(tmp_path, n=6)
    root = str(tmp_path)
    num_frames = 4
    sch = drivers.reenact_schedule(root, num_frames)
    img_list = sorted(glob.glob(root + '/*.png'))
    want_ids, want_cams = [], []
    for k, img_path in enumerate([img for img in img_list]):                      # the script's loop (:125-131, :159-160)
        if k > num_frames:
            break
        if k < 1:
            continue
        if k + 1 >= len(labels):
            continue
        want_ids.append(os.path.basename(img_list[k]).split('.')[0])
        camera_params = (np.array(labels[k - 1][1]) + np.array(labels[k][1]) + np.array(labels[k + 1][1])) / 3
        want_cams.append(torch.tensor(camera_params).unsqueeze(0).float())
    assert sch['ids'] == want_ids and sch['obj_paths'] == [root + f'/{i}.obj' for i in want_ids]
    assert sch['lms_paths'] == [root + f'/{i}_kpt2d.txt' for i in want_ids]
    assert torch.equal(sch['cams'], torch.cat(want_cams))


def test_frame_pack_round_trip(tmp_path):
    from next3d_b200 import inputs
    _make_drive_root(tmp_path, n=5)
    root = str(tmp_path)
    sch = drivers.reenact_schedule(root, 10)
    verts = np.stack([inputs.load_frame(o, l)[0].numpy() for o, l in zip(sch['obj_paths'], sch['lms_paths'])])
    path = str(tmp_path / 'clip.n3dpack')
    inputs.write_frame_pack(path, verts, sch['cams'].numpy(), ids=sch['ids'])
    pack = inputs.FramePack(path, pin=False)
    assert len(pack) == len(sch['ids']) and pack.ids == sch['ids']
    assert np.array_equal(pack.verts, verts) and np.array_equal(pack.cams, sch['cams'].numpy())
    frames = list(pack)
    assert frames[1].shape == (1, verts.shape[1], 3) and np.array_equal(frames[1][0].numpy(), verts[1])
