"""CPU: the drop-in TriPlaneGenerator keeps the reference's state-dict contract and API, and refuses to run on CPU."""
import pytest
import torch

from next3d_b200 import config, weights
from oracle import generator as og, ref_shim


def _tiny():
    from next3d_b200.triplane_next3d import TriPlaneGenerator
    cfg = config.tiny_config(512)
    sd = weights.make_state_dict(cfg, seed=2)
    G = TriPlaneGenerator.from_config(cfg, sd, device='cpu')
    return cfg, sd, G


def test_state_dict_names_and_shapes():
    cfg, sd, G = _tiny()
    got = G.state_dict()
    assert sorted(got.keys()) == sorted(s[0] for s in config.param_spec(cfg)) and len(got) == 674
    for k, v in sd.items():
        assert tuple(got[k].shape) == tuple(v.shape), k
        assert torch.equal(got[k].float(), v.float()), k


def test_mapping_matches_oracle():
    cfg, sd, G = _tiny()
    z, c_cond, _, _ = weights.demo_inputs(cfg, 3, seed=2)
    ws = G.mapping(z, c_cond, truncation_psi=0.7, truncation_cutoff=14)
    ref = og.mapping(sd, cfg, z, c_cond, 0.7, 14)
    assert ws.shape == (3, 28, 512)
    assert (ws - ref).abs().max().item() < 1e-5


def test_no_cpu_fallback():
    cfg, sd, G = _tiny()
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, 1, seed=2)
    ws = G.mapping(z, c_cond)
    with pytest.raises(RuntimeError, match='CUDA'):
        G.synthesis(ws, c_cam, v, noise_mode='const')


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference not present')
def test_reference_state_dict_round_trip():
    """The reference's own loader path: copy_params_and_buffers(G_ref, G_new, require_all=True) (gen_samples_next3d.py:153)."""
    cfg, sd, G = _tiny()
    G_ref = ref_shim.build_reference_generator(cfg)
    G_ref.load_state_dict(sd)
    from torch_utils import misc
    misc.copy_params_and_buffers(G_ref, G, require_all=True)
    G.load_state_dict(G_ref.state_dict())
    assert set(G.state_dict().keys()) == set(G_ref.state_dict().keys())
