"""`torch_utils.ops.conv2d_resample` on the tcgen05 implicit-GEMM kernel (reference: torch_utils/ops/conv2d_resample.py:48-143).

Same signature.  NCHW in / NCHW out like the reference; internally the input is converted to the engine's NHWC split-bf16
layout and the convolution runs as n3d_conv_gemm launches:
    up == 2 : 4 transposed-conv parity classes + n3d_fir_up_epilogue   (reference path :114-131)
    down == 2: n3d_fir_down_split + one stride-2 GEMM                  (:108-111)
    otherwise: one GEMM                                                (:134-136)
`groups > 1` (the reference's fused modulated conv uses groups = batch) is executed group by group.
Supported configurations = what the generator uses: 1x1 / 3x3 kernels, padding = k // 2, up/down in {1, 2}, the
[1,3,3,1] x [1,3,3,1] / 64 FIR; anything else raises (no silent fallback).  The fused engine (next3d_b200.engine) does not go
through this wrapper -- it keeps activations in NHWC split form between layers.
"""
import torch

from ... import _lib
from ... import kernels as K
from . import upfirdn2d
from .upfirdn2d import _parse_padding, _get_filter_size  # noqa: F401  (imported by reference code)


def _get_weight_shape(w):
    return [int(sz) for sz in w.shape]


def _check_fir(f):
    ref = upfirdn2d.setup_filter([1, 3, 3, 1], device=f.device)
    if tuple(f.shape) != (4, 4) or not torch.allclose(f, ref, atol=1e-7):
        raise RuntimeError('conv2d_resample: only the [1,3,3,1] (x) [1,3,3,1] / 64 resampling filter is supported by the sm_100a kernels')


def _one_group(x, w, f, up, down, flip_weight):
    """x [N,Cin,H,W] fp32, w [Cout,Cin,k,k] fp32 -> [N,Cout,H',W'] fp32."""
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    dev = x.device
    if up > 1:
        flip_weight = not flip_weight                                   # conv2d_resample.py:127 (transposed conv with flip = not flip_weight)
    if not flip_weight and k > 1:
        w = w.flip([2, 3])
    w_hi, w_lo = K.pack_conv_weight(w)
    xn = x.permute(0, 2, 3, 1).contiguous()
    if down == 2 and k == 3:
        sh, sw = (H + 2) // 2, (W + 2) // 2
        assert H % 2 == 0 and W % 2 == 0
        a_hi = torch.empty(4 * N, sh, sw, Cin, dtype=torch.bfloat16, device=dev)
        a_lo = torch.empty_like(a_hi)
        K.fir_down_split(xn, a_hi, a_lo)
        out = torch.empty(N, H // 2, W // 2, Cout, device=dev)
        K.conv_gemm(a_hi, a_lo, w_hi, w_lo, K.taps_stride2(), N, H // 2, W // 2, a_img_mul=N, out_f32=out, f32_cstride=Cout)
        return out.permute(0, 3, 1, 2)
    a_hi = torch.empty(N, H, W, Cin, dtype=torch.bfloat16, device=dev)
    a_lo = torch.empty_like(a_hi)
    K.modulate_split(xn, None, a_hi, a_lo)
    if up == 2 and k == 3:
        raw = torch.empty(N, 2 * H + 1, 2 * W + 1, Cout, device=dev)
        K.conv_transposed_gemm(a_hi, a_lo, w_hi, w_lo, N, H, W, raw)
        out = torch.empty(N, 2 * H, 2 * W, Cout, device=dev)
        K.fir_up_epilogue(raw, Cout, None, None, None, 1.0, 1.0, -1.0, out_f32=out, f32_cstride=Cout)
        return out.permute(0, 3, 1, 2)
    out = torch.empty(N, H, W, Cout, device=dev)
    K.conv_gemm(a_hi, a_lo, w_hi, w_lo, K.taps_conv3x3() if k == 3 else K.taps_conv1x1(), N, H, W, out_f32=out, f32_cstride=Cout)
    return out.permute(0, 3, 1, 2)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    _lib.require_cuda(x, 'conv2d_resample(x)')
    out_channels, in_channels_per_group, kh, kw = _get_weight_shape(w)
    px0, px1, py0, py1 = _parse_padding(padding)
    if kh != kw or kh not in (1, 3) or up not in (1, 2) or down not in (1, 2) or (up == 2 and down == 2):
        raise RuntimeError(f'conv2d_resample: kernel {kh}x{kw}, up={up}, down={down} is not supported by the sm_100a kernels')
    if not (px0 == px1 == py0 == py1 == kh // 2):
        raise RuntimeError('conv2d_resample: only padding == kernel_size // 2 is supported by the sm_100a kernels')
    if in_channels_per_group % 8 != 0:
        raise RuntimeError('conv2d_resample: input channels per group must be a multiple of 8 (TMA stride alignment)')
    if up > 1 or down > 1:
        assert f is not None, 'resampling needs the FIR filter'
        _check_fir(f.to(x.device))
    dtype = x.dtype
    xf, wf = x.float(), w.float()
    if kh == 1 and down == 2:                                            # fast path :96-99: downsample first, then 1x1 conv
        xf = upfirdn2d.downsample2d(xf, f.to(x.device))
        up = down = 1
    post_up = kh == 1 and up == 2                                        # fast path :102-105: 1x1 conv first, then upsample
    outs = []
    cin_g, cout_g = in_channels_per_group, out_channels // groups
    for g in range(groups):
        xg = xf[:, g * cin_g:(g + 1) * cin_g]
        wg = wf[g * cout_g:(g + 1) * cout_g]
        outs.append(_one_group(xg, wg, f, 1 if post_up else up, down, flip_weight))
    y = outs[0] if groups == 1 else torch.cat(outs, 1)
    if post_up:
        y = upfirdn2d.upsample2d(y.contiguous(), f.to(x.device))
    return y.contiguous().to(dtype)
