"""Generate golden fixtures by running the REAL reference (imported from /root/reference through
oracle/ref_shim.py) on seeded synthetic weights and inputs.  Container-side only:

    python tests/golden/make_golden.py

What the reference executes: its own TriPlaneGenerator.mapping + .synthesis on CPU (every torch_utils.ops call
takes its `_ref` branch, fp32), with (a) our seeded state dict loaded via load_state_dict, (b) the sampler's
uniforms injected, (c) our CPU restatement standing in for the absent pytorch3d rasterizer, (d) a synthetic
all-ones eye mask.  Outputs are stored subsampled so the fixtures stay small:
  image_raw, image_depth (full), image[..., ::4, ::4], face index buffer of sample 0 (int16, -1 = empty),
  alpha planes checksum, mouth boxes, and strided samples of the blended tri-planes (hooked from the renderer).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from next3d_b200 import config, weights  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = {
    # name: (config factory, img_resolution, batch, seed)
    'tiny512_b2': (config.tiny_config, 512, 2, 0),
    'tiny256_b1': (config.tiny_config, 256, 1, 1),      # BASELINE.json configs[0]: 64^2 -> 256^2 (4X SR head)
    'full512_b1': (config.full_config, 512, 1, 0),      # BASELINE.json configs[1] shapes, batch 1
}


def run_case(name):
    factory, res, batch, seed = CASES[name]
    cfg = factory(res)
    G = ref_shim.build_reference_generator(cfg)
    sd = weights.make_state_dict(cfg, seed=seed)
    G.load_state_dict(sd)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, batch, seed=seed)
    u_c, u_f = weights.sampler_noise(cfg, batch, seed=seed)
    captured = {}
    orig = G.renderer.forward

    def hook(planes, *a, **k):
        captured['planes'] = planes.detach().clone()
        return orig(planes, *a, **k)

    G.renderer.forward = hook
    with torch.no_grad():
        ws = G.mapping(z, c_cond, truncation_psi=0.7, truncation_cutoff=14)
        with ref_shim.injected_sampler_noise(u_c, u_f):
            out = G.synthesis(ws, c_cam, v, noise_mode='const')
    planes = captured['planes']
    np.savez_compressed(
        os.path.join(ROOT, 'tests', 'golden', f'{name}.npz'),
        ws=ws.numpy(),
        image_raw=out['image_raw'].numpy(),
        image_depth=out['image_depth'].numpy(),
        image_s4=out['image'][..., ::4, ::4].numpy(),
        image_absmax=np.float32(out['image'].abs().max().item()),
        image_sum=np.float64(out['image'].double().sum().item()),
        planes_s8=planes[..., ::8, ::8].numpy(),
        planes_sum=np.float64(planes.double().sum().item()),
    )
    print(name, 'image', tuple(out['image'].shape), 'range', out['image'].min().item(), out['image'].max().item())


if __name__ == '__main__':
    for n in (sys.argv[1:] or CASES):
        run_case(n)
