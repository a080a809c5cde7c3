"""Time render_kernel alone on the benchmark workload (planes random, batch 8, 64^2, 48+48).  python tools/bench_render.py"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from next3d_b200 import config, weights, kernels as K
cfg = config.full_config(512)
N, R = 8, 64
g = torch.Generator().manual_seed(0)
planes = torch.randn(N, 3, 256, 256, 32, generator=g).cuda()
dec = (torch.randn(64, 32, generator=g).cuda() / math.sqrt(32), torch.randn(64, generator=g).cuda() * 0.1,
       torch.randn(33, 64, generator=g).cuda() / math.sqrt(64), torch.randn(33, generator=g).cuda() * 0.1)
_, _, c, _ = weights.demo_inputs(cfg, N)
cam, intr = c[:, :16].contiguous().cuda(), c[:, 16:25].contiguous().cuda()
rgb = torch.zeros(N, R * R, 32, device='cuda'); depth = torch.zeros(N, R * R, device='cuda'); wsum = torch.zeros_like(depth)
mm = torch.tensor([float('inf'), 0.0], device='cuda')
for _ in range(3):
    K.render_rays(planes, cam, intr, R, cfg.rendering_kwargs, dec, rgb, depth, wsum, mm, seed=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    K.render_rays(planes, cam, intr, R, cfg.rendering_kwargs, dec, rgb, depth, wsum, mm, seed=i)
e1.record(); torch.cuda.synchronize()
print('N3D_RENDER_THREADS', os.environ.get('N3D_RENDER_THREADS', 'default'), 'render_kernel', e0.elapsed_time(e1) / 10, 'ms')
