"""CPU: the oracle restatement reproduces the committed golden fixtures (generated from the REAL reference by
tests/golden/make_golden.py).  Runs everywhere (no /root/reference needed)."""
import numpy as np
import pytest
import torch

from next3d_b200 import config, weights
from oracle import generator as og
from tests.golden.make_golden import CASES
from tests.helpers import load_golden, range_rel_err

# The reference and the oracle run the same fp32 torch ops; on the machine that produced the fixtures they agree
# bit for bit.  A different CPU / thread count may change summation order inside oneDNN/MKL: allow 2e-5 of range.
TOL = 2e-5


@pytest.mark.parametrize('name', ['tiny512_b2', 'tiny256_b1'])
def test_oracle_matches_golden(name):
    factory, res, batch, seed = CASES[name]
    cfg = factory(res)
    g = load_golden(name)
    sd = weights.make_state_dict(cfg, seed=seed)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, batch, seed=seed)
    u_c, u_f = weights.sampler_noise(cfg, batch, seed=seed)
    with torch.no_grad():
        ws = og.mapping(sd, cfg, z, c_cond, 0.7, 14)
        assert range_rel_err(ws, g['ws']) < 1e-6
        out = og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f, return_intermediates=True)
    assert range_rel_err(out['planes'][..., ::8, ::8], g['planes_s8']) < TOL
    assert range_rel_err(out['image_raw'], g['image_raw']) < TOL
    assert range_rel_err(out['image_depth'], g['image_depth']) < TOL
    assert range_rel_err(out['image'][..., ::4, ::4], g['image_s4']) < TOL
    assert abs(out['image'].abs().max().item() - float(g['image_absmax'])) < TOL * float(g['image_absmax'])
    # demo mesh: the landmark-derived mouth box
    assert out['mouth_boxes'].tolist() == [[82, 122, 108, 148]] * batch


def test_state_dict_contract():
    """674 tensors with the reference's names."""
    cfg = config.full_config(512)
    spec = config.param_spec(cfg)
    names = [s[0] for s in spec]
    assert len(names) == 674 and len(set(names)) == 674
    for must in ['texture_backbone.synthesis.b64.conv0.affine.weight', 'backbone.synthesis.b4.const',
                 'mouth_backbone.synthesis.encoder.3.conv2.bias', 'neural_blending.synthesis.fusion.2.weight',
                 'superresolution.block1.torgb.weight', 'decoder.net.2.bias', 'face_uvcoords', 'backbone.mapping.w_avg']:
        assert must in names
    nparam = sum(int(np.prod(s[1])) for s in spec if s[1] is not None and s[2] not in ('filter', 'noise_const', 'w_avg'))
    assert abs(nparam - 172.8e6) < 0.2e6
