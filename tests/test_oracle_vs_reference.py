"""CPU, container only: pin the oracle restatement against the reference's OWN code imported from
/root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from next3d_b200 import config, weights
from oracle import ops as oo, generator as og, rasterize as orast, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='/root/reference not present')


@pytest.fixture(scope='module')
def ref():
    ref_shim.import_reference()
    from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, filtered_lrelu
    import training_avatar_texture.networks_stylegan2 as sg2
    import training_avatar_texture.volumetric_rendering.renderer as rr
    return dict(bias_act=bias_act, upfirdn2d=upfirdn2d, conv2d_resample=conv2d_resample, filtered_lrelu=filtered_lrelu,
                sg2=sg2, rr=rr)


def test_bias_act_all_activations(ref):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 7, 5, 6, generator=g) * 3
    b = torch.randn(7, generator=g)
    for act in oo.ACTIVATIONS:
        for gain, clamp in [(None, None), (0.7, 0.5)]:
            a = ref['bias_act'].bias_act(x, b, act=act, gain=gain, clamp=clamp, impl='ref')
            o = oo.bias_act(x, b, act=act, gain=gain, clamp=clamp)
            assert torch.equal(a, o), act


@pytest.mark.parametrize('up,down,pad', [(1, 1, [1, 1, 1, 1]), (2, 1, [2, 1, 2, 1]), (1, 2, [1, 1, 1, 1]),
                                        (1, 1, [2, 2, 2, 2]), (2, 2, [3, 0, 1, 2]), (1, 1, [-1, 0, 0, -1])])
def test_upfirdn2d(ref, up, down, pad):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 9, 12, generator=g)
    f = ref['upfirdn2d'].setup_filter([1, 3, 3, 1])
    assert torch.equal(f, oo.setup_filter())
    a = ref['upfirdn2d'].upfirdn2d(x, f, up=up, down=down, padding=pad, gain=up * up, impl='ref')
    o = oo.upfirdn2d(x, oo.setup_filter(), up=up, down=down, padding=pad, gain=up * up)
    assert torch.equal(a, o)
    assert torch.equal(ref['upfirdn2d'].upsample2d(x, f, impl='ref'), oo.upsample2d(x, f))
    assert torch.equal(ref['upfirdn2d'].downsample2d(x[..., :8, :], f, impl='ref'), oo.downsample2d(x[..., :8, :], f))


@pytest.mark.parametrize('k,up,down,flip', [(3, 2, 1, False), (3, 1, 2, True), (3, 1, 1, True), (1, 1, 1, True)])
def test_conv2d_resample(ref, k, up, down, flip):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 6, 8, 8, generator=g)
    w = torch.randn(4, 6, k, k, generator=g)
    f = oo.setup_filter()
    a = ref['conv2d_resample'].conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=flip)
    o = oo.conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=flip)
    assert torch.equal(a, o)


def test_modulated_conv2d(ref):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 8, 8, generator=g)
    w = torch.randn(4, 6, 3, 3, generator=g)
    s = torch.randn(2, 6, generator=g)
    noise = torch.randn(16, 16, generator=g)
    f = oo.setup_filter()
    a = ref['sg2'].modulated_conv2d(x, w, s, noise=noise, up=2, padding=1, resample_filter=f, flip_weight=False)
    o = oo.modulated_conv2d(x, w, s, noise=noise, up=2, padding=1, resample_filter=f, flip_weight=False)
    assert torch.equal(a, o)


def test_filtered_lrelu(ref):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 10, 10, generator=g)
    b = torch.randn(3, generator=g)
    fu = ref['upfirdn2d'].setup_filter([1, 3, 3, 1])
    a = ref['filtered_lrelu'].filtered_lrelu(x, fu=fu, fd=fu, b=b, up=2, down=2, padding=3, impl='ref')
    o = oo.filtered_lrelu(x, fu=fu, fd=fu, b=b, up=2, down=2, padding=3)
    assert torch.equal(a, o)


def test_fill_mouth_vs_cv2(ref):
    """Our BFS restatement == the reference's cv2.floodFill-based fill_mouth, including fractional alphas."""
    g = torch.Generator().manual_seed(5)
    a = torch.zeros(3, 1, 64, 64)
    a[:, :, 10:50, 12:52] = 1.0
    a[0, :, 25:30, 20:40] = 0.0                      # hole
    a[1, :, 20:24, 20:24] = 0.5                      # interior fractional patch
    a[1, :, 40:44, 30:34] = 0.0
    a[2, :, :5, :5] = 1.0                            # corner covered
    a[2, :, 30:33, 30:36] = 0.0
    a = (a + torch.nn.functional.avg_pool2d(a, 3, 1, 1)) / 2   # fractional edges
    assert torch.equal(ref['rr'].fill_mouth(a), og.fill_mouth(a))


def test_full_synthesis_bit_exact(ref):
    cfg = config.tiny_config(512)
    G = ref_shim.build_reference_generator(cfg)
    sd = weights.make_state_dict(cfg, seed=3)
    G.load_state_dict(sd)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, 2, seed=3, jitter=1e-4)
    u_c, u_f = weights.sampler_noise(cfg, 2, seed=3)
    with torch.no_grad():
        ws_r = G.mapping(z, c_cond, truncation_psi=0.7, truncation_cutoff=14)
        ws = og.mapping(sd, cfg, z, c_cond, 0.7, 14)
        assert torch.equal(ws, ws_r)
        with ref_shim.injected_sampler_noise(u_c, u_f):
            r = G.synthesis(ws_r, c_cam, v, noise_mode='const')
        o = og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f)
    for k in ('image', 'image_raw', 'image_depth'):
        assert (o[k] - r[k]).abs().max().item() <= 1e-6 * r[k].abs().max().item(), k


def test_view_transform_is_the_blas_fma_chain(ref):
    """oracle.generator._bmm3_fma (explicit k = 0,1,2 fused chain, float64-emulated) == torch.bmm bit for bit: pins the summation
    order the reference's bmm (renderer.py:505-514) takes on this machine, which the CUDA transform kernel reproduces."""
    g = torch.Generator().manual_seed(0)
    p = torch.randn(3, 5023, 3, generator=g)
    R = torch.randn(3, 3, 3, generator=g)
    assert torch.equal(og._bmm3_fma(p, R), torch.bmm(p, R))
    for view in ((0, 0, 0), (0, 90, 0), (0, -90, 0), (90, 0, 0)):
        Rv = og.angle2matrix(torch.tensor(view, dtype=torch.float32))[None].expand(3, -1, -1)
        assert torch.equal(og._bmm3_fma(p, Rv), torch.bmm(p, Rv))
