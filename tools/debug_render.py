"""Print the fused renderer's errors against the oracle for a list of shapes (no asserts; development aid).
python tools/debug_render.py [N res Dc Df]..."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from next3d_b200 import config, weights, kernels as K
from oracle import renderer as orr

def rr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()

cases = [(3, 64, 48, 48), (2, 32, 48, 48), (1, 16, 96, 96), (1, 24, 36, 36), (1, 20, 48, 0), (1, 16, 40, 24), (2, 12, 12, 70), (2, 64, 96, 96)]
if len(sys.argv) > 1:
    a = [int(x) for x in sys.argv[1:]]
    cases = [tuple(a[i:i + 4]) for i in range(0, len(a), 4)]
DEV = 'cuda'
QUANT = int(os.environ.get('N3D_DEBUG_QUANT', '0'))
for N, res, D, Df in cases:
    cfg = config.tiny_config()
    opts = dict(cfg.rendering_kwargs, depth_resolution=D, depth_resolution_importance=Df)
    g = torch.Generator().manual_seed(47 + res)
    planes = torch.randn(N, 3, 32, 64, 64, generator=g)
    sd = {'decoder.net.0.weight': torch.randn(64, 32, generator=g), 'decoder.net.0.bias': torch.randn(64, generator=g) * 0.1,
          'decoder.net.2.weight': torch.randn(33, 64, generator=g), 'decoder.net.2.bias': torch.randn(33, generator=g) * 0.1}
    _, _, c, _ = weights.demo_inputs(cfg, N, seed=5)
    u_c = torch.rand(N, res * res, D, 1, generator=g)
    u_f = torch.rand(N * res * res, max(Df, 1), generator=g)
    if QUANT:
        u_f = torch.floor(u_f * QUANT) / QUANT + 0.01        # heavy ties and crowded buckets
    cam, intr = c[:, :16].reshape(-1, 4, 4), c[:, 16:25].reshape(-1, 3, 3)
    o, d = orr.ray_sampler(cam, intr, res)
    rgb_ref, depth_ref, w_ref = orr.render(sd, planes, o, d, opts, u_c, u_f)
    dec = ((sd['decoder.net.0.weight'] / math.sqrt(32)).to(DEV).contiguous(), sd['decoder.net.0.bias'].to(DEV),
           (sd['decoder.net.2.weight'] / math.sqrt(64)).to(DEV).contiguous(), sd['decoder.net.2.bias'].to(DEV))
    rgb = torch.zeros(N, res * res, 32, device=DEV); depth = torch.zeros(N, res * res, device=DEV); wsum = torch.zeros(N, res * res, device=DEV)
    mm = torch.tensor([float('inf'), 0.0], device=DEV)
    K.render_rays(planes.permute(0, 1, 3, 4, 2).contiguous().to(DEV), c[:, :16].contiguous().to(DEV), c[:, 16:25].contiguous().to(DEV), res, opts,
                  dec, rgb, depth, wsum, mm, u_coarse=u_c.to(DEV), u_fine=u_f.to(DEV))
    torch.cuda.synchronize()
    K.depth_clamp(depth, mm)
    e = (rgb.cpu() - rgb_ref).abs().amax(-1)
    print(f'case N={N} res={res} Dc={D} Df={Df}: rgb {rr(rgb.cpu(), rgb_ref):.2e} wsum {rr(wsum.cpu(), w_ref[..., 0]):.2e} '
          f'depth {rr(depth.cpu(), depth_ref[..., 0]):.2e} mm {mm.tolist()} bad_rays {(e > 1e-4).sum().item()}/{e.numel()} '
          f'nan {torch.isnan(rgb).sum().item()}', flush=True)
    if rr(wsum.cpu(), w_ref[..., 0]) > 1e-3:
        print('  wsum gpu', wsum.flatten()[:8].tolist(), '\n  wsum ref', w_ref[..., 0].flatten()[:8].tolist())
        print('  depth gpu', depth.flatten()[:8].tolist(), '\n  depth ref', depth_ref[..., 0].flatten()[:8].tolist())
