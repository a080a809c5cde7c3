"""Batched frame drivers: the per-frame Python loops of `gen_videos_next3d.py:128-171` and
`reenact_avatar_next3d.py:125-167` restated as (1) a frame schedule computed up front -- orbit cameras, interpolated latents --
and (2) a loop over BATCHES of frames through `TriPlaneGenerator.synthesis` (CUDA-graph replay), with the uint8 conversion on
the device and an asynchronous read-back, so that nothing but the generator sits on the critical path.

Everything here is host logic around the hot path; the per-frame values (camera matrices, interpolated ws, uint8 pixels) are the
reference's, checked in tests/test_drivers_cpu.py against the reference's own code.
"""
import math

import numpy as np
import torch

from . import camera

FOCAL_FFHQ = 4.2647            # gen_videos_next3d.py:97,139


def orbit_camera_params(num_frames, lookat, radius, yaw_range=0.35, pitch_range=0.25, focal=FOCAL_FFHQ):
    """-> c [num_frames, 25] float32: the camera sweep of gen_videos_next3d.py:128-140 (note the script's 3.14 literals)."""
    intr = torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=torch.float32)
    half = num_frames // 2
    out = []
    for f in range(num_frames):
        pose = camera.look_at_pose(3.14 / 2 + yaw_range * np.sin(2 * 3.14 * f / half),
                                   3.14 / 2 - 0.05 + pitch_range * np.cos(2 * 3.14 * f / half), lookat, radius)
        out.append(torch.cat([pose.reshape(16), intr.reshape(9)]))
    return torch.stack(out)


def interpolate_ws(ws_keyframes, w_frames, wraps=2, kind='cubic'):
    """ws_keyframes [K, L, D] (one grid cell of gen_videos_next3d.py:106-117) -> [K * w_frames, L, D]: the latents of every
    frame, `interp(frame_idx / w_frames)` for frame_idx = 0 .. K*w_frames-1 (:143-144), evaluated in one call."""
    import scipy.interpolate
    K = ws_keyframes.shape[0]
    x = np.arange(-K * wraps, K * (wraps + 1))
    y = np.tile(ws_keyframes.detach().cpu().numpy(), [wraps * 2 + 1, 1, 1])
    interp = scipy.interpolate.interp1d(x, y, kind=kind, axis=0)
    return torch.from_numpy(interp(np.arange(K * w_frames) / w_frames))


def to_uint8_hwc(img):
    """[N,3,H,W] float in [-1,1] -> [N,H,W,3] uint8, the arithmetic of layout_grid (gen_videos_next3d.py:40-46) for a 1x1 grid."""
    return (img * 127.5 + 128).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def render_frames(G, ws_frames, cams, verts, batch=8, image_mode='image', noise_mode='const', device=None, seed=None, sampler_noise=None):
    """Yield one uint8 HWC numpy frame per (ws, camera, mesh) triple, in order, rendering `batch` frames per synthesis call.

    ws_frames [F, L, D], cams [F, 25], verts: [F, V, 3] / [1, V, 3] (static mesh) tensors, or an iterable of per-frame [1, V, 3]
    tensors (e.g. inputs.FramePrefetcher).  The last, partial batch is padded by repeating its final frame (one graph shape) and
    the padding is dropped.  The device->host copy of batch i overlaps the synthesis of batch i+1.  `seed`: base seed of the ray
    sampler's RNG (batch k uses seed + k); None = a fresh random seed per call, like the reference's torch.rand.
    `sampler_noise=(u_coarse [F, M, Dc, 1], u_fine [F, M, Df])`: per-frame sampler uniforms instead of the in-kernel RNG (parity tests:
    frame f gets the same noise whichever batch slot it lands in)."""
    device = device or next(G.parameters()).device
    F = ws_frames.shape[0]
    cuda = torch.device(device).type == 'cuda'
    vit = None
    if not torch.is_tensor(verts):
        vit = iter(verts)
    copy_s = torch.cuda.Stream(device) if cuda else None
    pending = None                                               # (host tensor, event, valid count) of the previous batch

    def drain(p):
        host, ev, n = p
        if ev is not None:
            ev.synchronize()
        for k in range(n):
            yield host[k].numpy()

    for b0 in range(0, F, batch):
        n = min(batch, F - b0)
        idx = list(range(b0, b0 + n)) + [b0 + n - 1] * (batch - n)
        w = ws_frames[idx].to(device, torch.float32, non_blocking=True)
        c = cams[idx].to(device, torch.float32, non_blocking=True)
        if vit is not None:
            vs = [next(vit) for _ in range(n)]
            v = torch.cat(vs + [vs[-1]] * (batch - n), 0).to(device, torch.float32, non_blocking=True)
        elif verts.shape[0] == 1:
            v = verts.to(device, torch.float32).expand(batch, -1, -1)
        else:
            v = verts[idx].to(device, torch.float32, non_blocking=True)
        kw = {} if seed is None else {'seed': int(seed) + b0 // batch}
        if sampler_noise is not None:
            u_c, u_f = sampler_noise
            kw['sampler_noise'] = (u_c[idx], u_f[idx].reshape(-1, u_f.shape[-1]))
        img = G.synthesis(w, c, v, noise_mode=noise_mode, **kw)[image_mode]
        if image_mode == 'image_depth':                          # gen_videos_next3d.py:160-162, per frame
            img = -img
            lo, hi = img.amin(dim=(1, 2, 3), keepdim=True), img.amax(dim=(1, 2, 3), keepdim=True)
            img = (img - lo) / (hi - lo) * 2 - 1
            img = img.expand(-1, 3, -1, -1) if img.shape[1] == 1 else img
        u8 = to_uint8_hwc(img)                                   # fresh tensor: the generator's static output buffer is free again
        if cuda:
            host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
            copy_s.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(copy_s):
                host.copy_(u8, non_blocking=True)
                u8.record_stream(copy_s)
                ev = torch.cuda.Event()
                ev.record(copy_s)
        else:
            host, ev = u8, None
        if pending is not None:
            yield from drain(pending)
        pending = (host, ev, n)
    if pending is not None:
        yield from drain(pending)


# ------------------------------------------------------------------------------------------------ reenactment (reenact_avatar_next3d.py)
def reenact_schedule(drive_root, num_frames, lms_cond=True):
    """Frame schedule of reenact_avatar_next3d.py:125-160 for a driving directory (dataset.json + NNN.png / NNN.obj / NNN_kpt2d.txt):
    the script walks the sorted PNG list, skips k = 0, stops after k = num_frames and smooths the camera over the labels k-1, k, k+1
    (so the last usable frame is len(labels) - 2; the script itself runs into an IndexError there).
    -> dict(ks, ids, obj_paths, lms_paths (or None), cams float32 [F, 25])."""
    import glob
    import os
    from . import inputs
    labels = inputs.load_labels(os.path.join(drive_root, 'dataset.json'))
    img_list = sorted(glob.glob(drive_root + '/*.png'))
    ks = [k for k in range(len(img_list)) if 1 <= k <= num_frames and k + 1 < len(labels)]
    ids = [os.path.basename(img_list[k]).split('.')[0] for k in ks]
    return dict(ks=ks, ids=ids, obj_paths=[drive_root + f'/{i}.obj' for i in ids],
                lms_paths=[drive_root + f'/{i}_kpt2d.txt' for i in ids] if lms_cond else None,
                cams=inputs.smoothed_cameras(labels, ks))


def reenact_frames(G, ws, drive_root, num_frames, batch=8, lms_cond=True, fixed_camera=None, noise_mode='const', seed=None,
                   sampler_noise=None, prefetch_depth=8, workers=4):
    """reenact_avatar_next3d.py:125-167 for ONE identity `ws` [1, L, D]: yields the rendered uint8 HWC frame of every driving frame,
    in order.  Meshes are parsed by the native parsers on worker threads (inputs.FramePrefetcher, pinned memory) while the GPU renders
    batches of `batch` frames; `fixed_camera` [1, 25] replaces the per-frame cameras (the script's --fixed_camera).
    A `.n3dpack` file written by inputs.write_frame_pack can be given instead of a directory (no parsing at all)."""
    from . import inputs
    if str(drive_root).endswith('.n3dpack'):
        pack = inputs.FramePack(drive_root)
        n = min(len(pack), num_frames)
        cams, verts = torch.from_numpy(np.array(pack.cams[:n])), (t for i, t in enumerate(pack) if i < n)
    else:
        sch = reenact_schedule(drive_root, num_frames, lms_cond)
        n = len(sch['ks'])
        cams = sch['cams']
        pairs = list(zip(sch['obj_paths'], sch['lms_paths'] if lms_cond else [None] * n))
        verts = inputs.FramePrefetcher(pairs, depth=prefetch_depth, workers=workers)
    if fixed_camera is not None:
        cams = fixed_camera.reshape(1, 25).cpu().float().expand(n, -1)
    if n == 0:
        return
    wsf = ws.detach().cpu().float().reshape(1, *ws.shape[-2:]).expand(n, -1, -1)
    yield from render_frames(G, wsf, cams, verts, batch=batch, noise_mode=noise_mode, seed=seed, sampler_noise=sampler_noise)


def interpolate_ws_device(ws_keyframes, w_frames, wraps=2, kind='cubic'):
    """interpolate_ws on the device: the spline is linear in its knots, so its basis B [K * w_frames, K] (interp1d evaluated once on
    the identity, wrap-around tiling folded in) is built on the host and the frames are B @ ws_keyframes in one small kernel
    (n3d_interp_rows) -- the latents never visit the host.  float32; equals interpolate_ws to ~1e-6."""
    import scipy.interpolate
    from . import kernels as K
    Kf = ws_keyframes.shape[0]
    x = np.arange(-Kf * wraps, Kf * (wraps + 1))
    eye = np.tile(np.eye(Kf), [wraps * 2 + 1, 1])                                  # knot i of the tiled sequence is keyframe i % K
    B = scipy.interpolate.interp1d(x, eye, kind=kind, axis=0)(np.arange(Kf * w_frames) / w_frames)
    B = torch.from_numpy(B).to(ws_keyframes.device, torch.float32).contiguous()
    return K.interp_rows(B, ws_keyframes.to(torch.float32).contiguous())


def render_frames_sharded(G, ws_frames, cams, verts, batch=8, noise_mode='const', seed=0, device=None, out=None):
    """Strong-scaling video driver (BASELINE.json configs[3]: one clip split over the GPUs of a box): rank r renders the contiguous
    frame range shard_range(F, r, world) in batches through `G.synthesis`, converts to uint8 HWC on the device (4x fewer bytes than
    the fp32 images) and the batches are gathered to rank 0 over NCCL on a side stream from a ping-pong staging buffer, so batch
    i+1 renders while batch i is in flight; rank 0 copies every gathered batch to pinned host memory on a third stream.
    Returns the [F, H, W, 3] uint8 host tensor (frame order) on rank 0, None elsewhere.  World size 1: no collective.
    `batch` is an upper bound: the per-call batch is the size in [batch/2, batch] that pads the rank's frame count least (30 frames per
    rank -> 5 batches of 6, not 4 of 8).  `out`: optional pinned [F, H, W, 3] uint8 host tensor to fill on rank 0 (pinning 190 MB per
    clip costs tens of milliseconds)."""
    import torch.distributed as dist
    from . import distributed as D
    device = device or next(G.parameters()).device
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    F = ws_frames.shape[0]
    spans = [D.shard_range(F, r, world) for r in range(world)]
    a, b = spans[rank]
    nmax = max(e - s for s, e in spans)
    batch = min(range(max(1, batch // 2), batch + 1), key=lambda bb: (-(-nmax // bb) * bb, -bb))     # least padding, then the largest
    steps = -(-nmax // batch)                                             # every rank runs the same number of collective steps
    res = G.img_resolution
    host = None
    if rank == 0:
        host = out if out is not None else torch.empty(F, res, res, 3, dtype=torch.uint8).pin_memory()
    main = torch.cuda.current_stream(device)
    side, copy_s = torch.cuda.Stream(device), torch.cuda.Stream(device)
    stage = [torch.empty(batch, res, res, 3, dtype=torch.uint8, device=device) for _ in range(2)]
    gath = [torch.empty(world, batch, res, res, 3, dtype=torch.uint8, device=device) if rank == 0 and world > 1 else None for _ in range(2)]
    drained = [None, None]
    static_mesh = verts.shape[0] == 1
    for k in range(steps):
        lo = min(a + k * batch, max(b - 1, a))
        idx = [min(lo + j, b - 1) for j in range(batch)]                   # the tail repeats the last frame (one graph shape)
        w = ws_frames[idx].to(device, torch.float32, non_blocking=True)
        c = cams[idx].to(device, torch.float32, non_blocking=True)
        v = verts.to(device, torch.float32).expand(batch, -1, -1) if static_mesh else verts[idx].to(device, torch.float32, non_blocking=True)
        img = G.synthesis(w, c, v, noise_mode=noise_mode, seed=int(seed) + rank * steps + k)['image']
        s = k & 1
        if drained[s] is not None:
            main.wait_event(drained[s])                                    # batch k-2 has left this staging buffer
        stage[s].copy_(to_uint8_hwc(img))
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if world > 1:
                dist.gather(stage[s], list(gath[s].unbind(0)) if rank == 0 else None, dst=0)
            ev = torch.cuda.Event()
            ev.record(side)
        drained[s] = ev
        if rank == 0:
            copy_s.wait_event(ev)
            with torch.cuda.stream(copy_s):
                for r in range(world):
                    s0, e0 = spans[r]
                    f0 = s0 + k * batch
                    n = max(0, min(batch, e0 - f0))
                    if n > 0:
                        src = gath[s][r] if world > 1 else stage[s]
                        host[f0:f0 + n].copy_(src[:n], non_blocking=True)
                ev2 = torch.cuda.Event()
                ev2.record(copy_s)
            drained[s] = ev2                                               # the staging / gather buffers are free once the D2H left them
    main.wait_stream(side)
    main.wait_stream(copy_s)
    return host


# ------------------------------------------------------------------------------------------------ shape extraction (row f2)
def create_samples(N, cube_length, head=0, count=None, device='cpu'):
    """Voxel-grid query points [1, count, 3] of `create_samples` (gen_samples_next3d.py:80-102) for the flat indices
    head .. head+count-1, computed with the reference's own float32 operations (including its float -- not floor -- division,
    which shears the y / x columns by idx / N, and the float32 rounding of indices above 2^24), so that any chunk is bit-identical
    to the corresponding slice of the reference's full N^3 tensor without materialising it (201 MB at N = 256, 1.6 GB at 512)."""
    count = N ** 3 - head if count is None else count
    voxel_origin = np.array([0, 0, 0]) - cube_length / 2
    voxel_size = cube_length / (N - 1)
    idx = torch.arange(head, head + count, dtype=torch.int64, device=device)
    s = torch.zeros(count, 3, device=device)
    s[:, 2] = idx % N
    s[:, 1] = (idx.float() / N) % N
    s[:, 0] = ((idx.float() / N) / N) % N
    s[:, 0] = (s[:, 0] * voxel_size) + voxel_origin[2]
    s[:, 1] = (s[:, 1] * voxel_size) + voxel_origin[1]
    s[:, 2] = (s[:, 2] * voxel_size) + voxel_origin[0]
    return s.unsqueeze(0)


def trim_sigma_grid(sigmas, shape_res, pad_value=-1000.0):
    """flip + border trim of gen_samples_next3d.py:226-238 on a [R,R,R] tensor (device or host)."""
    sigmas = torch.flip(sigmas, dims=[0])
    pad = int(30 * shape_res / 256)
    for d in range(3):
        sl = [slice(None)] * 3
        sl[d] = slice(0, pad)
        sigmas[tuple(sl)] = pad_value
        sl[d] = slice(shape_res - pad, shape_res)
        sigmas[tuple(sl)] = pad_value
    return sigmas


def extract_sigma_grid(G, ws, v, shape_res=512, max_batch=1000000, noise_mode='const'):
    """Density grid for marching cubes (gen_samples_next3d.py:208-238) -> float32 numpy [R,R,R], flipped and trimmed like the
    script's.  The tri-planes are computed ONCE (the script's G.sample re-runs the three backbones for every 1M-point chunk: 17
    times at 256^3, 135 times at 512^3), the query points are generated inside the decoding kernel and only sigma is decoded."""
    from . import kernels as K
    eng = G._get_engine()
    eng.rk = G.rendering_kwargs
    planes = eng.compute_planes(ws, v, noise_mode)
    box = G.rendering_kwargs['box_warp']
    sigmas = torch.empty(shape_res, shape_res, shape_res, device=planes.device)
    # one launch: voxel centres generated in the kernel (bit-identical to create_samples), flip + border trim fused into the store,
    # border voxels (55 % of a 256^3 grid) never decoded.  max_batch is kept for API compatibility with the script's chunk size.
    K.sample_grid(planes[0], shape_res, box * 1, box, eng.dec, sigmas, pad=int(30 * shape_res / 256))
    return sigmas.cpu().numpy()
