"""GPU check of drivers.render_frames: batched + graph-replayed frames == one-frame-at-a-time synthesis (tiny config)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from next3d_b200 import config, weights, drivers
from next3d_b200.triplane_next3d import TriPlaneGenerator
cfg = config.tiny_config(512)
G = TriPlaneGenerator.from_config(cfg, weights.make_state_dict(cfg, 3), device='cuda')
z, cc, c, v = weights.demo_inputs(cfg, 2, seed=3)
ws = G.mapping(z.cuda(), cc.cuda(), truncation_psi=0.7, truncation_cutoff=14)
F = 5
wsf = drivers.interpolate_ws(ws.cpu(), 3, wraps=1)[:F]
cams = drivers.orbit_camera_params(F, torch.tensor([0, 0, 0.2]), 2.7)
G.use_cuda_graph = True
torch.manual_seed(0)
got = list(drivers.render_frames(G, wsf, cams, v[:1], batch=2))
G.use_cuda_graph = False
worst = 0
for f in range(F):
    ref = drivers.to_uint8_hwc(G.synthesis(wsf[f:f + 1].cuda().float(), cams[f:f + 1].cuda(), v[:1].cuda(), noise_mode='const', seed=0)['image'])[0].cpu().numpy()
    worst = max(worst, int(np.abs(got[f].astype(int) - ref.astype(int)).max()))
print('frames', len(got), got[0].shape, got[0].dtype, 'max |uint8 diff| vs per-frame (different sampler seeds):', worst)
