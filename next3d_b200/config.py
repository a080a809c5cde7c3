"""Static description of the Next3D generator (shapes, parameter names, rendering options).

This is the *contract* between the reference's pickles / state dicts and this implementation: every
tensor name and shape below mirrors the reference module tree (SURVEY.md section 8b "state-dict contract"):
  TriPlaneGenerator.__init__            training_avatar_texture/triplane_next3d.py:41-109
  StyleGAN2 Generator / SynthesisNetwork  training_avatar_texture/networks_stylegan2.py:596-682
  StyleUNet Generator / SynthesisNetwork  training_avatar_texture/networks_stylegan2_styleunet.py:494-625
  Superresolution modules               training_avatar_texture/superresolution.py:62-88, 264-290
  OSGDecoder                            training_avatar_texture/triplane_next3d.py:348-357
"""
import copy
import dataclasses
import math


FFHQ_RENDERING_KWARGS = dict(  # train_next3d.py:313-339 (cfg == 'ffhq')
    image_resolution=512, disparity_space_sampling=False, clamp_mode='softplus',
    c_gen_conditioning_zero=False, gpc_reg_prob=0.5, c_scale=1.0,
    superresolution_noise_mode='none', density_reg=0.25, density_reg_p_dist=0.004, reg_type='l1',
    decoder_lr_mul=1.0, sr_antialias=True, gen_exp_cond=False,
    depth_resolution=48, depth_resolution_importance=48, ray_start=2.25, ray_end=3.3, box_warp=1,
    avg_camera_radius=2.7, avg_camera_pivot=[0, 0, 0.2],
)


@dataclasses.dataclass
class GeneratorConfig:
    img_resolution: int = 512          # 512 -> SuperresolutionHybrid8XDC, 256 -> ...4X
    channel_base: int = 32768
    channel_max: int = 512
    neural_rendering_resolution: int = 64
    rendering_kwargs: dict = dataclasses.field(default_factory=lambda: copy.deepcopy(FFHQ_RENDERING_KWARGS))
    w_dim: int = 512
    z_dim: int = 512
    c_dim: int = 25
    plane_res: int = 256               # resolution of texture / static / blended planes
    plane_ch: int = 32
    sr_num_fp16_res: int = 4           # train_next3d.py:196 default; > 0 => SR blocks clamp at +-256 (superresolution.py:271-277)

    @property
    def sr_module(self):
        return {512: '8XDC', 256: '4X'}[self.img_resolution]

    @property
    def sr_clamp(self):
        return 256.0 if self.sr_num_fp16_res > 0 else None

    def channels(self, res):
        return min(self.channel_base // res, self.channel_max)


def full_config(img_resolution=512):
    return GeneratorConfig(img_resolution=img_resolution)


def tiny_config(img_resolution=512):
    """Same topology, 16x fewer backbone channels (channel_max 32; SR keeps its hard-coded widths): the
    oracle runs in about a second, used by parity tests."""
    return GeneratorConfig(img_resolution=img_resolution, channel_base=4096, channel_max=32)


# ----------------------------------------------------------------------------------------------
# Parameter / buffer specification.  kind in {'weight', 'bias', 'affine_bias', 'noise_const',
# 'noise_strength', 'const', 'filter', 'w_avg', 'topology'}.

def _fc(prefix, out_f, in_f, spec, bias_kind='bias', lr_mul=1.0):
    spec.append((f'{prefix}.weight', (out_f, in_f), 'weight', dict(lr_mul=lr_mul)))
    spec.append((f'{prefix}.bias', (out_f,), bias_kind, {}))


def _mapping(prefix, cfg, spec):
    # MappingNetwork(num_layers=2): networks_stylegan2.py:193-231
    _fc(f'{prefix}.embed', cfg.w_dim, cfg.c_dim, spec)
    _fc(f'{prefix}.fc0', cfg.w_dim, cfg.z_dim + cfg.w_dim, spec, lr_mul=0.01)
    _fc(f'{prefix}.fc1', cfg.w_dim, cfg.w_dim, spec, lr_mul=0.01)
    spec.append((f'{prefix}.w_avg', (cfg.w_dim,), 'w_avg', {}))


def _synth_layer(prefix, cin, cout, res, cfg, spec, k=3):
    # SynthesisLayer: networks_stylegan2.py:276-309
    spec.append((f'{prefix}.resample_filter', (4, 4), 'filter', {}))
    spec.append((f'{prefix}.noise_const', (res, res), 'noise_const', {}))
    _fc(f'{prefix}.affine', cin, cfg.w_dim, spec, bias_kind='affine_bias')
    spec.append((f'{prefix}.weight', (cout, cin, k, k), 'weight', {}))
    spec.append((f'{prefix}.noise_strength', (), 'noise_strength', {}))
    spec.append((f'{prefix}.bias', (cout,), 'bias', {}))


def _torgb(prefix, cin, cimg, cfg, spec):
    # ToRGBLayer: networks_stylegan2.py:340-351
    _fc(f'{prefix}.affine', cin, cfg.w_dim, spec, bias_kind='affine_bias')
    spec.append((f'{prefix}.weight', (cimg, cin, 1, 1), 'weight', {}))
    spec.append((f'{prefix}.bias', (cimg,), 'bias', {}))


def _synth_block(prefix, cin, cout, res, cimg, cfg, spec):
    # SynthesisBlock (architecture='skip'): networks_stylegan2.py:492-542
    spec.append((f'{prefix}.resample_filter', (4, 4), 'filter', {}))
    if cin == 0:
        spec.append((f'{prefix}.const', (cout, res, res), 'const', {}))
    else:
        _synth_layer(f'{prefix}.conv0', cin, cout, res, cfg, spec)
    _synth_layer(f'{prefix}.conv1', cout, cout, res, cfg, spec)
    _torgb(f'{prefix}.torgb', cout, cimg, cfg, spec)


def block_resolutions(img_res):
    return [2 ** i for i in range(2, int(math.log2(img_res)) + 1)]


def _synthesis_network(prefix, cfg, cimg, spec):
    for res in block_resolutions(cfg.plane_res):
        cin = cfg.channels(res // 2) if res > 4 else 0
        _synth_block(f'{prefix}.b{res}', cin, cfg.channels(res), res, cimg, cfg, spec)


def _conv2d_layer(prefix, cin, cout, k, bias, spec):
    # styleunet Conv2dLayer: networks_stylegan2_styleunet.py:159-196
    spec.append((f'{prefix}.resample_filter', (4, 4), 'filter', {}))
    spec.append((f'{prefix}.weight', (cout, cin, k, k), 'weight', {}))
    if bias:
        spec.append((f'{prefix}.bias', (cout,), 'bias', {}))


def encoder_resolutions(in_size, final_size):
    return [2 ** i for i in range(int(math.log2(in_size)), int(math.log2(final_size)) - 1, -1)]


def _styleunet(prefix, cfg, in_size, final_size, spec):
    # styleunet SynthesisNetwork.__init__: networks_stylegan2_styleunet.py:494-552
    _synthesis_network(prefix, cfg, cfg.plane_ch, spec)
    enc = encoder_resolutions(in_size, final_size)
    for i, res in enumerate(enc[:-1]):
        cin, cout = cfg.channels(res), cfg.channels(res // 2)
        p = f'{prefix}.encoder.{i}'
        _conv2d_layer(f'{p}.fromrgb', cfg.plane_ch, cin, 1, False, spec)
        _conv2d_layer(f'{p}.conv1', cin, cin, 3, True, spec)
        _conv2d_layer(f'{p}.conv2', cin, cout, 3, True, spec)
        spec.append((f'{p}.resample_filter', (4, 4), 'filter', {}))
    for i, res in enumerate(enc[::-1]):
        c = cfg.channels(res)
        _conv2d_layer(f'{prefix}.fusion.{i}', c * 2 if res > final_size else c, c, 3, True, spec)


def sr_channels(cfg):
    # SuperresolutionHybrid8XDC: 32 -> 256 -> 128 (superresolution.py:273-276); 4X: 32 -> 128 -> 64 (:71-74).
    # The reference hard-codes these (channel_base/channel_max are ignored, superresolution.py:267).
    return {'8XDC': (256, 128), '4X': (128, 64)}[cfg.sr_module]


def param_spec(cfg):
    """Ordered list of (name, shape, kind, extra) for the whole TriPlaneGenerator state dict."""
    spec = []
    # texture_backbone (triplane_next3d.py:63)
    _synthesis_network('texture_backbone.synthesis', cfg, cfg.plane_ch, spec)
    _mapping('texture_backbone.mapping', cfg, spec)
    # mouth_backbone (triplane_next3d.py:64): in_size 64, final_size 4
    _styleunet('mouth_backbone.synthesis', cfg, 64, 4, spec)
    _mapping('mouth_backbone.mapping', cfg, spec)
    # backbone (triplane_next3d.py:65): 96 channels, mapping_ws = 28
    _synthesis_network('backbone.synthesis', cfg, cfg.plane_ch * 3, spec)
    _mapping('backbone.mapping', cfg, spec)
    # superresolution (triplane_next3d.py:67)
    c0, c1 = sr_channels(cfg)
    if cfg.sr_module == '8XDC':
        _synth_block('superresolution.block0', cfg.plane_ch, c0, 256, 3, cfg, spec)
        _synth_block('superresolution.block1', c0, c1, 512, 3, cfg, spec)
    else:
        spec.append(('superresolution.resample_filter', (4, 4), 'filter', {}))
        _synth_block('superresolution.block0', cfg.plane_ch, c0, 128, 3, cfg, spec)  # SynthesisBlockNoUp
        _synth_block('superresolution.block1', c0, c1, 256, 3, cfg, spec)
    # decoder (triplane_next3d.py:348-357)
    _fc('decoder.net.0', 64, cfg.plane_ch, spec)
    _fc('decoder.net.2', 1 + 32, 64, spec)
    # topology buffers (triplane_next3d.py:85-103)
    for name in ('dense_faces', 'faces', 'raw_uvcoords', 'uvcoords', 'uvfaces', 'face_uvcoords'):
        spec.append((name, None, 'topology', {}))
    # neural_blending (triplane_next3d.py:109): in_size 256, final_size 32
    _styleunet('neural_blending.synthesis', cfg, 256, 32, spec)
    _mapping('neural_blending.mapping', cfg, spec)
    return spec
