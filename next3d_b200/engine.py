"""Generator engine: packs a reference state dict once and runs `TriPlaneGenerator.synthesis` as a fixed sequence of
libnext3d_b200 kernel launches (no ATen compute on the hot path; torch only owns the device buffers and the stream).

Execution model:
  * all style vectors and demodulation coefficients of the ~70 modulated layers: 2 launches (n3d_styles, n3d_demod);
  * every convolution: n3d_conv_gemm (tcgen05 implicit GEMM, shared weights for the whole batch -- the algebraically
    equivalent non-fused form of networks_stylegan2.py:70-79: activations are pre-multiplied by the consumer's style,
    demodulation is applied in the epilogue), with noise / bias / lrelu / clamp / next-layer modulation / bf16 split fused;
  * up-convs: 4 parity-class GEMMs (stride-2 transposed conv, conv2d_resample.py:114-127) + one FIR-epilogue kernel;
  * down-convs: FIR + parity split kernel, then one GEMM over 9 unit-stride taps;
  * mesh path, volume renderer, resizes, blending: one kernel each (raster.cu, renderer.cu).
Reference call order: triplane_next3d.py:117-188.
"""
import math
import os

import numpy as np
import torch

from . import config as _config
from . import kernels as K

SQRT2 = math.sqrt(2.0)
VIEWS = ((0, 0, 0), (0, 90, 0), (0, -90, 0), (90, 0, 0))      # triplane_next3d.py:140-145


def _angle2matrix(angles_deg):
    """fp32 sin/cos of angle*pi/180, rows [cz*cy, ...] (volumetric_rendering/renderer.py:518-547)."""
    a = torch.tensor(angles_deg, dtype=torch.float32).reshape(1, 3) * np.pi / 180.
    s, c = torch.sin(a), torch.cos(a)
    cx, cy, cz = c[0, 0], c[0, 1], c[0, 2]
    sx, sy, sz = s[0, 0], s[0, 1], s[0, 2]
    return torch.stack([cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                        sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                        -sy, cy * sx, cy * cx]).reshape(3, 3)


class Split:
    """bf16 (hi, lo) NHWC activation pair."""
    __slots__ = ('hi', 'lo')

    def __init__(self, shape, device):
        self.hi = torch.empty(shape, dtype=torch.bfloat16, device=device)
        self.lo = torch.empty(shape, dtype=torch.bfloat16, device=device)

    @property
    def C(self):
        return self.hi.shape[-1]


class _ModLayer:
    """A modulated layer (SynthesisLayer or ToRGBLayer) after packing."""

    def __init__(self, name, cin, cout, k, up, widx, is_rgb, clamp):
        self.name, self.cin, self.cout, self.k, self.up, self.widx, self.is_rgb, self.clamp = name, cin, cout, k, up, widx, is_rgb, clamp
        self.w_hi = self.w_lo = self.bias = self.noise = None
        self.sbase = self.dbase = -1          # element offsets (per N=1) of this layer's style / dcoef blocks


class _PlainLayer:
    """A non-modulated Conv2dLayer (StyleUNet encoder / fusion)."""

    def __init__(self, name, cin, cout, k):
        self.name, self.cin, self.cout, self.k = name, cin, cout, k
        self.w_hi = self.w_lo = self.bias = None


class Engine:
    def __init__(self, cfg, state_dict, device='cuda', nprod=3, uv_face_mask=None):
        if not torch.cuda.is_available():
            raise RuntimeError('next3d_b200.Engine needs a CUDA device (sm_100a); there is no CPU path')
        self.cfg, self.device, self.nprod = cfg, torch.device(device), nprod
        self.rk = cfg.rendering_kwargs
        self.mod, self.plain = {}, {}
        self.launches = 0
        self.conv_flops = 0.0
        self.prof = None
        self._cur_layer = ''
        sd = {k: v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for k, v in state_dict.items()}
        self._pack(sd)
        self._tables = {}
        mask = torch.ones(1, 1, 256, 256) if uv_face_mask is None else uv_face_mask
        self.eye_mask = mask.reshape(mask.shape[-2], mask.shape[-1]).to(self.device, torch.float32).contiguous()
        self._mm_init = torch.tensor([float('inf'), 0.0], device=self.device)
        self._graphs = {}
        self._side = None                       # side streams of the concurrent branches (compute_planes)
        self.concurrent = os.environ.get('N3D_CONCURRENT', '1') != '0'
        self.splitk = os.environ.get('N3D_SPLITK', '1') != '0'
        self.splitk_maxres = int(os.environ.get('N3D_SPLITK_MAXRES', '16'))

    # ------------------------------------------------------------------------------------------ packing (one-time)
    def _add_mod(self, sd, name, cin, cout, k, up, widx, is_rgb=False, clamp=None, noise=True):
        L = _ModLayer(name, cin, cout, k, up, widx, is_rgb, clamp)
        w = sd[f'{name}.weight'].float()
        L.w_hi, L.w_lo = K.pack_conv_weight(w)
        L.w_f32 = w.reshape(cout, cin).contiguous() if (is_rgb and cout <= 4) else None    # fused-ToRGB epilogue operand
        L.bias = sd[f'{name}.bias'].float().contiguous()
        L.aff_w = sd[f'{name}.affine.weight'].float().contiguous()
        L.aff_b = sd[f'{name}.affine.bias'].float().contiguous()
        L.wsq = None if is_rgb else w.square().sum(dim=[2, 3]).contiguous()        # [Cout, Cin]
        if not is_rgb and noise:
            L.noise = (sd[f'{name}.noise_const'].float() * sd[f'{name}.noise_strength'].float()).contiguous()
            L.noise_strength = sd[f'{name}.noise_strength'].float()
        self.mod[name] = L
        return L

    def _add_plain(self, sd, name, k):
        w = sd[f'{name}.weight'].float()
        cout, cin = w.shape[:2]
        L = _PlainLayer(name, cin, cout, k)
        L.w_hi, L.w_lo = K.pack_conv_weight(w, gain=1.0 / math.sqrt(cin * k * k))       # Conv2dLayer.weight_gain
        b = sd.get(f'{name}.bias')
        L.bias = b.float().contiguous() if b is not None else None
        self.plain[name] = L
        return L

    def _pack_synthesis(self, sd, prefix, cimg, ws_off, first_res=4):
        cfg = self.cfg
        for res in _config.block_resolutions(cfg.plane_res):
            k = int(math.log2(res)) - 2
            start = 0 if res == 4 else 1 + 2 * (k - 1)
            cin = cfg.channels(res // 2) if res > 4 else 0
            cout = cfg.channels(res)
            p = f'{prefix}.b{res}'
            if res < first_res:
                continue                                   # e.g. mouth_backbone.b4 / neural_blending.b4..b32: never executed
            if res > 4:
                self._add_mod(sd, f'{p}.conv0', cin, cout, 3, 2, ws_off + start)
                self._add_mod(sd, f'{p}.conv1', cout, cout, 3, 1, ws_off + start + 1)
                self._add_mod(sd, f'{p}.torgb', cout, cimg, 1, 1, ws_off + start + 2, is_rgb=True)
            else:
                self._add_mod(sd, f'{p}.conv1', cout, cout, 3, 1, ws_off + 0)
                self._add_mod(sd, f'{p}.torgb', cout, cimg, 1, 1, ws_off + 1, is_rgb=True)
                self.consts = getattr(self, 'consts', {})
                self.consts[prefix] = sd[f'{p}.const'].float().permute(1, 2, 0).contiguous()      # [4,4,C] NHWC

    def _pack_unet(self, sd, prefix, in_size, final_size):
        enc = _config.encoder_resolutions(in_size, final_size)
        for i, _ in enumerate(enc[:-1]):
            self._add_plain(sd, f'{prefix}.encoder.{i}.fromrgb', 1)
            self._add_plain(sd, f'{prefix}.encoder.{i}.conv1', 3)
            self._add_plain(sd, f'{prefix}.encoder.{i}.conv2', 3)
        n_fusion = len(enc) - 1                            # the last fusion module (highest res) is never executed
        for i in range(n_fusion):
            self._add_plain(sd, f'{prefix}.fusion.{i}', 3)
        self._pack_synthesis(sd, prefix, self.cfg.plane_ch, 0, first_res=final_size * 2)

    def _pack(self, sd):
        cfg = self.cfg
        self._pack_synthesis(sd, 'texture_backbone.synthesis', cfg.plane_ch, 14)
        self._pack_synthesis(sd, 'backbone.synthesis', cfg.plane_ch * 3, 0)
        self._pack_unet(sd, 'mouth_backbone.synthesis', 64, 4)
        self._pack_unet(sd, 'neural_blending.synthesis', 256, 32)
        c0, c1 = _config.sr_channels(cfg)
        sr_up0 = 2 if cfg.sr_module == '8XDC' else 1
        for blk, (cin, cout, up) in enumerate([(cfg.plane_ch, c0, sr_up0), (c0, c1, 2)]):
            p = f'superresolution.block{blk}'
            noise = self.rk['superresolution_noise_mode'] != 'none'
            # ws = eg3d_ws[:, -1:].repeat(1, 3, 1) (superresolution.py:280): every SR layer reads ws index 13
            self._add_mod(sd, f'{p}.conv0', cin, cout, 3, up, 13, clamp=cfg.sr_clamp, noise=noise)
            self._add_mod(sd, f'{p}.conv1', cout, cout, 3, 1, 13, clamp=cfg.sr_clamp, noise=noise)
            self._add_mod(sd, f'{p}.torgb', cout, 3, 1, 1, 13, is_rgb=True, clamp=cfg.sr_clamp)
        # decoder (OSGDecoder, triplane_next3d.py:348-357): FullyConnectedLayer gains folded like its forward does
        self.dec = ((sd['decoder.net.0.weight'].float() * (1.0 / math.sqrt(32))).contiguous(), sd['decoder.net.0.bias'].float().contiguous(),
                    (sd['decoder.net.2.weight'].float() * (1.0 / math.sqrt(64))).contiguous(), sd['decoder.net.2.bias'].float().contiguous())
        # mapping network stays in PyTorch (adjacent to, not inside, synthesis)
        self.mapping_sd = {k[len('backbone.mapping.'):]: v.float() for k, v in sd.items() if k.startswith('backbone.mapping.')}
        # topology
        self.faces = sd['faces'][0][:, [0, 2, 1]].to(torch.int32).contiguous()                       # :207
        self.face_uv = sd['face_uvcoords'][0][:, [0, 2, 1], :2].float().contiguous()                # :208 (u, v)
        self.rot = torch.stack([_angle2matrix(a) for a in VIEWS]).to(self.device).contiguous()
        # concatenated affine tables (layer order fixed here)
        order = list(self.mod.values())
        sb = db = 0
        for L in order:
            L.sbase = sb
            sb += L.cin
            if not L.is_rgb:
                L.dbase = db
                db += L.cout
        self.style_elems, self.dcoef_elems = sb, db
        self.aff_w = torch.cat([L.aff_w for L in order], 0).contiguous()
        self.aff_b = torch.cat([L.aff_b for L in order], 0).contiguous()
        self.row_widx = torch.cat([torch.full((L.cin,), L.widx, dtype=torch.int32) for L in order]).to(self.device)
        self.row_scale = torch.cat([torch.full((L.cin,), (1.0 / math.sqrt(L.cin * L.k * L.k)) if L.is_rgb else 1.0) for L in order]).to(self.device)
        self.row_cin = torch.cat([torch.full((L.cin,), L.cin, dtype=torch.int32) for L in order]).to(self.device)
        dl = [L for L in order if not L.is_rgb]
        self.wsq = torch.cat([L.wsq.reshape(-1) for L in dl]).contiguous()
        woff, acc = [], 0
        for L in dl:
            woff.append(torch.arange(L.cout, dtype=torch.int64) * L.cin + acc)
            acc += L.cout * L.cin
        self.d_woff = torch.cat(woff).to(self.device)
        self.d_cin = torch.cat([torch.full((L.cout,), L.cin, dtype=torch.int32) for L in dl]).to(self.device)
        self.d_cout = torch.cat([torch.full((L.cout,), L.cout, dtype=torch.int32) for L in dl]).to(self.device)
        for L in order:
            del L.aff_w, L.aff_b
            L.wsq = None

    def _batch_tables(self, N):
        """Per-batch-size offset tables: every layer's style / dcoef block is dense [N, C] at base*N."""
        t = self._tables.get(N)
        if t is None:
            order = list(self.mod.values())
            s_ooff = torch.cat([torch.arange(L.cin, dtype=torch.int64) + L.sbase * N for L in order]).to(self.device)
            dl = [L for L in order if not L.is_rgb]
            d_soff = torch.cat([torch.full((L.cout,), L.sbase * N, dtype=torch.int64) for L in dl]).to(self.device)
            d_ooff = torch.cat([torch.arange(L.cout, dtype=torch.int64) + L.dbase * N for L in dl]).to(self.device)
            t = self._tables[N] = (s_ooff, d_soff, d_ooff)
        return t

    # ------------------------------------------------------------------------------------------ instrumentation
    def _prof_begin(self):
        if self.prof is None:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def _prof_end(self, e0, kind, flops=0, info=None):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.prof.append((kind, e0, e1, flops, info))

    def _gemm(self, a_hi, a_lo, w_hi, w_lo, taps, N, MH, MW, **kw):
        """n3d_conv_gemm + bookkeeping: algorithmic FLOPs = 2 * Cin * Cout * taps * M-space positions (one product)."""
        flops = 2.0 * a_hi.shape[-1] * w_hi.shape[1] * len(taps) * N * MH * MW
        self.conv_flops += flops
        ev = self._prof_begin()
        K.conv_gemm(a_hi, a_lo, w_hi, w_lo, taps, N, MH, MW, **kw)
        self._prof_end(ev, 'conv_gemm', flops, (self._cur_layer, a_hi.shape[-1], w_hi.shape[1], MH, MW, len(taps)))

    def profile_summary(self):
        """After a synthesis() with self.prof = []: {kind: (launches, total_ms, total_flops)} (synchronises)."""
        torch.cuda.synchronize(self.device)
        out = {}
        for kind, e0, e1, fl, _ in self.prof or []:
            n, ms, f = out.get(kind, (0, 0.0, 0.0))
            out[kind] = (n + 1, ms + e0.elapsed_time(e1), f + fl)
        return out

    # ------------------------------------------------------------------------------------------ small helpers
    def _style(self, L):
        return self._styles[L.sbase * self._N: (L.sbase + L.cin) * self._N]

    def _dcoef(self, L):
        return self._dcoefs[L.dbase * self._N: (L.dbase + L.cout) * self._N]

    def _f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def _noise(self, L, noise_mode):
        if L.noise is None or noise_mode == 'none':
            return None, 0
        if noise_mode == 'const':
            return L.noise, 0
        res = L.noise.shape[0]                               # 'random': per-sample noise (networks_stylegan2.py:318-319)
        return (torch.randn(self._N, res, res, device=self.device) * L.noise_strength).contiguous(), res * res

    def _out(self, buf, consumer=None, coff=0):
        """SplitOut descriptor writing into `buf` pre-multiplied by `consumer`'s style (None = unmodulated)."""
        style = self._style(self.mod[consumer]) if consumer else None
        return K.make_split_out(buf.hi, buf.lo, style, buf.C, coff)

    # ------------------------------------------------------------------------------------------ layer runners
    def _fused_rgb(self, name, img, accumulate, nchw=False):
        """Descriptor that folds ToRGBLayer `name` (<= 4 image channels) into the epilogue of the conv producing its input."""
        L = self.mod[name]
        return dict(out=img, weight=L.w_f32, style=self._style(L), bias=L.bias, clamp=L.clamp if L.clamp is not None else -1.0,
                    nchw=nchw, accumulate=accumulate)

    # ---- split-K for the 4^2 / 8^2 layers: a batch of 8 gives 1-4 tiles of 128 pixels and a 72-step serial K loop (9 taps x 512
    # channels): the nine taps become nine concurrent work items per tile writing raw partial sums, summed in fixed order by
    # n3d_splitk_epilogue (deterministic; an atomics-based variant was rejected for run-to-run differences).
    def _use_splitk(self, res):
        return self.splitk and res <= self.splitk_maxres   # by resolution only: the summation order of a sample must not depend on the batch size

    def _splitk_conv(self, name, a, L, res, epi):
        N, S = self._N, 9
        part = self._f32(S, N, res, res, L.cout)
        self._gemm(a.hi, a.lo, L.w_hi, L.w_lo, K.taps_conv3x3(), N, res, res, nprod=self.nprod, mode=1, out_f32=part, f32_cstride=L.cout,
                    splits=S, split_stride=N * res * res * L.cout)
        ev = self._prof_begin()                  # the reduction pass is part of this convolution: counted as conv time (0 extra FLOPs)
        K.splitk_epilogue(part, **epi)
        self._prof_end(ev, 'conv_gemm', 0.0, (name, a.hi.shape[-1], L.cout, res, res, 0))
        self.launches += 2

    def _modconv(self, name, a, res_in, outs, noise_mode, f32=None, rgb=None):
        """SynthesisLayer (networks_stylegan2.py:311-330); `a` already carries this layer's modulation."""
        L, N = self.mod[name], self._N
        self._cur_layer = name
        noise, nstride = self._noise(L, noise_mode)
        clamp = L.clamp if L.clamp is not None else -1.0
        if L.up == 1 and self._use_splitk(res_in) and rgb is None:
            self._splitk_conv(name, a, L, res_in, dict(dcoef=self._dcoef(L), bias=L.bias, noise=noise, noise_nstride=nstride, gain=SQRT2, slope=0.2,
                                                        clamp=clamp, outs=outs, out_f32=f32, f32_cstride=L.cout if f32 is not None else 0))
            return
        if L.up == 1:
            self._gemm(a.hi, a.lo, L.w_hi, L.w_lo, K.taps_conv3x3(), N, res_in, res_in, nprod=self.nprod, dcoef=self._dcoef(L), bias=L.bias,
                        noise=noise, noise_nstride=nstride, gain=SQRT2, slope=0.2, clamp=clamp, outs=outs, out_f32=f32,
                        f32_cstride=L.cout if f32 is not None else 0, rgb=rgb)
            self.launches += 1
            return
        assert rgb is None
        raw = self._f32(N, 2 * res_in + 1, 2 * res_in + 1, L.cout)
        flops = 2.0 * L.cin * L.cout * N * (3 * res_in + 2) ** 2          # taps x positions summed over the 4 parity classes
        self.conv_flops += flops
        ev = self._prof_begin()
        K.conv_transposed_gemm(a.hi, a.lo, L.w_hi, L.w_lo, N, res_in, res_in, raw, nprod=self.nprod)
        self._prof_end(ev, 'conv_gemm', flops, (name, L.cin, L.cout, res_in + 1, res_in + 1, 9))
        K.fir_up_epilogue(raw, L.cout, self._dcoef(L), L.bias, noise, SQRT2, 0.2, clamp, outs=outs, out_f32=f32,
                          f32_cstride=L.cout if f32 is not None else 0, noise_nstride=nstride)
        self.launches += 2

    def _torgb(self, name, a, res, img, accumulate, nchw=False):
        """ToRGBLayer (networks_stylegan2.py:353-357): 1x1 modulated conv without demodulation, linear bias (+ clamp)."""
        L = self.mod[name]
        self._cur_layer = name
        self._gemm(a.hi, a.lo, L.w_hi, L.w_lo, K.taps_conv1x1(), self._N, res, res, nprod=self.nprod, bias=L.bias,
                    clamp=L.clamp if L.clamp is not None else -1.0, out_f32=img, f32_cstride=L.cout, f32_nchw=nchw, f32_accumulate=accumulate)
        self.launches += 1

    def _plainconv(self, name, a, res, act, outs=(), f32=None, accumulate=False, stride2=False):
        """styleunet Conv2dLayer (networks_stylegan2_styleunet.py:198-207): weight_gain folded into the packed weights."""
        L, N = self.plain[name], self._N
        self._cur_layer = name
        gain, slope = (SQRT2, 0.2) if act == 'lrelu' else (1.0, 1.0)
        if stride2:
            self._gemm(a.hi, a.lo, L.w_hi, L.w_lo, K.taps_stride2(), N, res // 2, res // 2, a_img_mul=N, nprod=self.nprod, bias=L.bias, gain=gain,
                        slope=slope, outs=outs, out_f32=f32, f32_cstride=L.cout if f32 is not None else 0, f32_accumulate=accumulate)
        elif L.k == 3 and self._use_splitk(res) and not accumulate:
            self._splitk_conv(name, a, L, res, dict(dcoef=None, bias=L.bias, noise=None, gain=gain, slope=slope, clamp=-1.0, outs=outs, out_f32=f32,
                                                    f32_cstride=L.cout if f32 is not None else 0))
            return
        else:
            taps = K.taps_conv3x3() if L.k == 3 else K.taps_conv1x1()
            self._gemm(a.hi, a.lo, L.w_hi, L.w_lo, taps, N, res, res, nprod=self.nprod, bias=L.bias, gain=gain, slope=slope, outs=outs,
                        out_f32=f32, f32_cstride=L.cout if f32 is not None else 0, f32_accumulate=accumulate)
        self.launches += 1

    # ------------------------------------------------------------------------------------------ networks
    def _synthesis_blocks(self, prefix, x, res_list, cimg, noise_mode, concat_next=None, final_nchw=None, img=None):
        """Run SynthesisBlocks `res_list` (all with conv0 up) starting from split activation `x` (modulated for the first
        conv0) at resolution res_list[0]//2.  concat_next: {res: (concat Split buffer, fusion runner)} -- when the output of
        block `res` feeds a fusion conv (StyleUNet) instead of the next conv0 directly."""
        N, dev = self._N, self.device
        for i, res in enumerate(res_list):
            p = f'{prefix}.b{res}'
            c = self.mod[f'{p}.conv1'].cout
            x1 = Split((N, res, res, c), dev)
            self._modconv(f'{p}.conv0', x, res // 2, [self._out(x1, f'{p}.conv1')], noise_mode)
            nxt = res_list[i + 1] if i + 1 < len(res_list) else None
            rgb_in = Split((N, res, res, c), dev)
            outs = [self._out(rgb_in, f'{p}.torgb')]
            x = None
            if concat_next and res in concat_next:
                cat_buf, _ = concat_next[res]
                outs.append(K.make_split_out(cat_buf.hi, cat_buf.lo, None, cat_buf.C, 0))
            elif nxt is not None:
                x = Split((N, res, res, c), dev)
                outs.append(self._out(x, f'{prefix}.b{nxt}.conv0'))
            self._modconv(f'{p}.conv1', x1, res, outs, noise_mode)
            last = nxt is None
            if img is None:
                img = self._f32(N, res, res, cimg)
                self._torgb(f'{p}.torgb', rgb_in, res, img, accumulate=False)
            else:
                if last and final_nchw is not None:
                    K.upsample2d_nhwc(img, final_nchw, y_nchw=True)
                    self._torgb(f'{p}.torgb', rgb_in, res, final_nchw, accumulate=True, nchw=True)
                    img = final_nchw
                else:
                    up = self._f32(N, res, res, cimg)
                    K.upsample2d_nhwc(img, up)
                    self._torgb(f'{p}.torgb', rgb_in, res, up, accumulate=True)
                    img = up
                self.launches += 1
            if concat_next and res in concat_next:
                x = concat_next[res][1]()                      # run the fusion conv, returns the next conv0's input
        return img

    def _backbone(self, prefix, cimg, noise_mode):
        """StyleGAN2 SynthesisNetwork (networks_stylegan2.py:630-645), 4 -> 256."""
        N, dev, cfg = self._N, self.device, self.cfg
        c4 = cfg.channels(4)
        const = self.consts[prefix][None].expand(N, -1, -1, -1).contiguous()
        x0 = Split((N, 4, 4, c4), dev)
        K.modulate_split(const, self._style(self.mod[f'{prefix}.b4.conv1']), x0.hi, x0.lo)
        rgb_in = Split((N, 4, 4, c4), dev)
        x = Split((N, 4, 4, c4), dev)
        self._modconv(f'{prefix}.b4.conv1', x0, 4, [self._out(rgb_in, f'{prefix}.b4.torgb'), self._out(x, f'{prefix}.b8.conv0')], noise_mode)
        img = self._f32(N, 4, 4, cimg)
        self._torgb(f'{prefix}.b4.torgb', rgb_in, 4, img, accumulate=False)
        self.launches += 1
        return self._synthesis_blocks(prefix, x, _config.block_resolutions(cfg.plane_res)[1:], cimg, noise_mode, img=img)

    def _styleunet(self, prefix, x_in, in_size, final_size, num_cond_res, noise_mode):
        """StyleUNet SynthesisNetwork.forward (networks_stylegan2_styleunet.py:554-588).  x_in: fp32 NHWC [N,in,in,32]."""
        N, dev, cfg = self._N, self.device, self.cfg
        enc_res = _config.encoder_resolutions(in_size, final_size)
        res_list = _config.block_resolutions(cfg.plane_res)
        start = int(math.log2(final_size)) - 1
        dec_res = res_list[start:]
        # which decoder levels get a condition: index j (block dec_res[j]) iff 2**(j + log2(final)) < num_cond_res
        n_cond = sum(1 for j in range(len(dec_res)) if 2 ** (j + int(math.log2(final_size))) < num_cond_res)
        # fusion j consumes cond_list[j] at resolution final_size * 2**j; for j >= 1 it is concatenated after the decoder's x
        cat = {}
        for j in range(1, n_cond):
            r = final_size * 2 ** j
            cat[j] = Split((N, r, r, 2 * cfg.channels(r)), dev)
        cond0 = Split((N, final_size, final_size, cfg.channels(final_size)), dev)
        # ---- encoder (EncoderResBlock.forward :107-115)
        skip = None
        x = x_in
        n_enc = len(enc_res) - 1
        for i, r in enumerate(enc_res[:-1]):
            p = f'{prefix}.encoder.{i}'
            cin = cfg.channels(r)
            if i > 0:
                xd = self._f32(N, r, r, cfg.plane_ch)
                K.downsample2d_nhwc(x, xd)
                self.launches += 1
                x = xd
            xs = Split((N, r, r, cfg.plane_ch), dev)
            K.modulate_split(x, None, xs.hi, xs.lo)
            self.launches += 1
            h = Split((N, r, r, cin), dev)
            if skip is None:
                self._plainconv(f'{p}.fromrgb', xs, r, 'linear', outs=[K.make_split_out(h.hi, h.lo, None, cin, 0)])
            else:
                self._plainconv(f'{p}.fromrgb', xs, r, 'linear', f32=skip, accumulate=True)       # out = fromrgb(x) + skip
                K.modulate_split(skip, None, h.hi, h.lo)
                self.launches += 1
            c1 = self._f32(N, r, r, cin)
            self._plainconv(f'{p}.conv1', h, r, 'lrelu', f32=c1)
            sh = (r + 2) // 2
            par = Split((4 * N, sh, sh, cin), dev)
            K.fir_down_split(c1, par.hi, par.lo)
            self.launches += 1
            # cond of this block lives at resolution r/2; cond_list (reversed) index:
            j = n_enc - 1 - i
            cout = cfg.channels(r // 2)
            outs = []
            if j == 0:
                outs.append(K.make_split_out(cond0.hi, cond0.lo, None, cond0.C, 0))
            elif j < n_cond:
                outs.append(K.make_split_out(cat[j].hi, cat[j].lo, None, cat[j].C, cat[j].C - cout))
            skip = self._f32(N, r // 2, r // 2, cout) if i + 1 < n_enc else None
            self._plainconv(f'{p}.conv2', par, r, 'lrelu', outs=outs, f32=skip, stride2=True)
        # ---- decoder with fusion
        first = dec_res[0]
        x = Split((N, final_size, final_size, cfg.channels(final_size)), dev)
        self._plainconv(f'{prefix}.fusion.0', cond0, final_size, 'linear', outs=[self._out(x, f'{prefix}.b{first}.conv0')])
        concat_next = {}
        for j in range(1, n_cond):
            r = final_size * 2 ** j

            def run_fusion(j=j, r=r):
                y = Split((N, r, r, cfg.channels(r)), dev)
                self._plainconv(f'{prefix}.fusion.{j}', cat[j], r, 'linear', outs=[self._out(y, f'{prefix}.b{2 * r}.conv0')])
                return y
            concat_next[r] = (cat[j], run_fusion)
        return self._synthesis_blocks(prefix, x, dec_res, cfg.plane_ch, noise_mode, concat_next=concat_next)

    def _superresolution(self, feat, noise_mode_sr, image_out):
        """SuperresolutionHybrid8XDC / 4X (superresolution.py:77-88, 279-290).  feat: fp32 NHWC [N,R,R,32]."""
        N, dev, cfg = self._N, self.device, self.cfg
        R = feat.shape[1]
        p0, p1 = 'superresolution.block0', 'superresolution.block1'
        c0, c1 = _config.sr_channels(cfg)
        need = (R != 128) if cfg.sr_module == '8XDC' else (R < 128)
        r0 = 128 if need else R
        if need and R > 128 and not self.rk.get('sr_antialias', True):
            raise RuntimeError('next3d_b200: sr_antialias=False with a neural rendering resolution above 128 (a plain bilinear minification) is not '
                               'implemented')
        x = Split((N, r0, r0, cfg.plane_ch), dev)
        rgb_lo = feat[..., :3].contiguous()                                    # image_raw channels (tiny copy)
        if need:
            K.resize_aa(feat, None, r0, r0, style=self._style(self.mod[f'{p0}.conv0']), hi=x.hi, lo=x.lo)
            rgb = self._f32(N, r0, r0, 3)
            K.resize_aa(rgb_lo, rgb)
            self.launches += 2
        else:
            K.modulate_split(feat, self._style(self.mod[f'{p0}.conv0']), x.hi, x.lo)
            rgb = rgb_lo
            self.launches += 1
        # block0
        res0 = r0 * (2 if cfg.sr_module == '8XDC' else 1)
        x1 = Split((N, res0, res0, c0), dev)
        self._modconv(f'{p0}.conv0', x, r0, [self._out(x1, f'{p0}.conv1')], noise_mode_sr)
        x2 = Split((N, res0, res0, c0), dev)
        if cfg.sr_module == '8XDC':
            img = self._f32(N, res0, res0, 3)
            K.upsample2d_nhwc(rgb, img)
            self.launches += 1
        else:
            img = rgb.clone()                                                   # SynthesisBlockNoUp: no image upsample
        # the two ToRGB layers (3 channels) run inside the conv1 epilogues: img += torgb(conv1 output)
        self._modconv(f'{p0}.conv1', x1, res0, [self._out(x2, f'{p1}.conv0')], noise_mode_sr, rgb=self._fused_rgb(f'{p0}.torgb', img, True))
        # block1
        res1 = res0 * 2
        x3 = Split((N, res1, res1, c1), dev)
        self._modconv(f'{p1}.conv0', x2, res0, [self._out(x3, f'{p1}.conv1')], noise_mode_sr)
        K.upsample2d_nhwc(img, image_out, y_nchw=True)
        self._modconv(f'{p1}.conv1', x3, res1, [], noise_mode_sr, rgb=self._fused_rgb(f'{p1}.torgb', image_out, True, nchw=True))
        self.launches += 1

    # ------------------------------------------------------------------------------------------ planes + synthesis
    def _on_device(self, **tensors):
        """Inputs must live on the engine's device: raw device pointers are handed to the C ABI, so a CPU or other-GPU tensor
        would become an illegal address instead of a clean error."""
        for name, t in tensors.items():
            if t is not None and t.device != self.device:
                raise RuntimeError(f'next3d_b200: `{name}` is on {t.device} but the generator runs on {self.device}; move it there '
                                   f'(there is no CPU path)')

    def compute_planes(self, ws, v, noise_mode='const', return_intermediates=False):
        """Everything up to the blended tri-planes (triplane_next3d.py:137-174) -> [N,3,256,256,32] channels-last fp32."""
        self._on_device(ws=ws, v=v)
        with torch.cuda.device(self.device):
            return self._compute_planes(ws, v, noise_mode, return_intermediates)

    def _compute_planes(self, ws, v, noise_mode='const', return_intermediates=False):
        cfg, dev = self.cfg, self.device
        N = ws.shape[0]
        self._N = N
        P = cfg.plane_res
        ws = ws.to(torch.float32).contiguous()
        v = v.to(torch.float32)
        verts, lms = v[:, :5023].contiguous(), v[:, 5023:].contiguous()
        s_ooff, d_soff, d_ooff = self._batch_tables(N)
        self._styles = self._f32(self.style_elems * N)
        self._dcoefs = self._f32(self.dcoef_elems * N)
        K.styles(ws, self.aff_w, self.aff_b, self.row_widx, self.row_scale, s_ooff, self.row_cin, self._styles)
        K.demod(self._styles, self.wsq, self.d_woff, self.d_cin, d_soff, d_ooff, self.d_cout, self._dcoefs, N)
        self.launches += 2
        # Three independent branches (triplane_next3d.py:137-170): (1) neural texture -> UV lookup -> mouth / blending UNets,
        # (2) mesh rasterization + mouth box (needs only the vertices), (3) the static tri-plane backbone (needed only by the final
        # blend).  (2) and (3) run on side streams: their latency-bound kernels (4^2..32^2 layers on a few dozen SMs, the flood
        # fill, the rasterizer) fill SMs the main branch leaves idle and vice versa.  Captured as parallel graph branches.
        main = torch.cuda.current_stream(dev)
        if self.concurrent:
            if self._side is None:
                self._side = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_rast, s_static = self._side
            s_rast.wait_stream(main)
            s_static.wait_stream(main)
        else:
            s_rast = s_static = main
        with torch.cuda.stream(s_rast):
            tv = self._f32(N, 4, verts.shape[1], 3)
            K.transform_points(verts, self.rot, 10.0, True, tv)
            tl = self._f32(N, 4, lms.shape[1], 3)
            K.transform_points(lms, self.rot, 0.0, False, tl)
            p2f = torch.empty(N * 4, P, P, dtype=torch.int32, device=dev)
            bary = self._f32(N * 4, P, P, 3)
            K.rasterize(tv.view(N * 4, -1, 3), self.faces, P, P, p2f, bary)
            boxes = torch.empty(N, 4, dtype=torch.int32, device=dev)
            lm2d = tl[:, 0, :, :2].contiguous()
            K.mouth_box(lm2d, boxes)
        with torch.cuda.stream(s_static):
            static = self._backbone('backbone.synthesis', cfg.plane_ch * 3, noise_mode)             # [N,256,256,96]
        textures = self._backbone('texture_backbone.synthesis', cfg.plane_ch, noise_mode)          # [N,256,256,32]
        if self.concurrent:
            main.wait_stream(s_rast)
            for t in (p2f, bary, boxes, lm2d):
                t.record_stream(main)
        # look up texture + eye mask at the rasterized UVs, fill the mouth hole
        tex_planes = self._f32(3, N, P, P, cfg.plane_ch)
        alpha = self._f32(3, N, P, P)
        K.uv_sample(p2f, bary, self.face_uv, textures, self.eye_mask, tex_planes, alpha)
        K.fill_mouth(alpha)
        self.launches += 7                                                          # rasterize = setup pass + bin pass
        # mouth crop -> StyleUNet -> paste back -> neural blending
        front = tex_planes[0]
        crop = self._f32(N, 64, 64, cfg.plane_ch)
        K.resize_aa(front, crop, src_box=boxes)
        mouth = self._styleunet('mouth_backbone.synthesis', crop, 64, 4, 64, noise_mode)
        stitched = front.clone()
        K.resize_aa(mouth, stitched, dst_box=boxes)
        blended = self._styleunet('neural_blending.synthesis', stitched, 256, 32, 256, noise_mode)
        if self.concurrent:
            main.wait_stream(s_static)
            static.record_stream(main)
        planes = self._f32(N, 3, P, P, cfg.plane_ch)
        K.blend_planes(blended, tex_planes, alpha, static, planes)
        self.launches += 3
        if return_intermediates:
            return planes, dict(textures=textures, pix_to_face=p2f.view(N, 4, P, P), tex_planes=tex_planes, alpha=alpha, boxes=boxes,
                                mouth_crop=crop, mouth_plane=mouth, stitched=stitched, blended_front=blended, static=static, lm2d=lm2d)
        return planes

    def synthesis(self, ws, c, v, noise_mode='const', neural_rendering_resolution=None, sampler_noise=None, seed=0,
                  return_intermediates=False, seed_ptr=None):
        self._on_device(ws=ws, c=c, v=v)
        with torch.cuda.device(self.device):
            return self._synthesis(ws, c, v, noise_mode, neural_rendering_resolution, sampler_noise, seed, return_intermediates, seed_ptr)

    def _synthesis(self, ws, c, v, noise_mode='const', neural_rendering_resolution=None, sampler_noise=None, seed=0,
                   return_intermediates=False, seed_ptr=None):
        cfg, dev = self.cfg, self.device
        N = ws.shape[0]
        R = neural_rendering_resolution or cfg.neural_rendering_resolution
        self.launches = 0
        self.conv_flops = 0.0
        out = self._compute_planes(ws, v, noise_mode, return_intermediates)
        planes, inter = out if return_intermediates else (out, None)
        c = c.to(torch.float32)
        cam = c[:, :16].contiguous()
        intr = c[:, 16:25].contiguous()
        M = R * R
        feat = self._f32(N, R, R, cfg.plane_ch)                                 # [N, M, 32] == NHWC feature image
        depth = self._f32(N, 1, R, R)
        wsum = self._f32(N, M)
        mm = self._mm_init.clone()                                          # running (min, max) of all sample depths
        u_c, u_f = sampler_noise if sampler_noise is not None else (None, None)
        if u_c is not None:
            u_c, u_f = u_c.to(dev, torch.float32).contiguous(), u_f.to(dev, torch.float32).contiguous()
        ev = self._prof_begin()
        K.render_rays(planes, cam, intr, R, self.rk, self.dec, feat, depth, wsum, mm, u_coarse=u_c, u_fine=u_f, seed=seed, seed_ptr=seed_ptr)
        self._prof_end(ev, 'render_rays')
        K.depth_clamp(depth, mm)
        image = self._f32(N, 3, cfg.img_resolution, cfg.img_resolution)
        self._superresolution(feat, self.rk['superresolution_noise_mode'], image)
        self.launches += 2
        result = {'image': image, 'image_raw': feat.permute(0, 3, 1, 2)[:, :3], 'image_depth': depth}
        if return_intermediates:
            inter.update(planes=planes, feature_image=feat, weights_sum=wsum)
            result['intermediates'] = inter
        return result

    # ------------------------------------------------------------------------------------------ CUDA graph replay
    def synthesis_graphed(self, ws, c, v, noise_mode='const', neural_rendering_resolution=None, seed=0):
        with torch.cuda.device(self.device):
            return self._synthesis_graphed(ws, c, v, noise_mode, neural_rendering_resolution, seed)

    def _synthesis_graphed(self, ws, c, v, noise_mode='const', neural_rendering_resolution=None, seed=0):
        """Same as synthesis() (in-kernel sampler RNG) but the ~200 launches of one forward are captured once per
        (batch, resolution, noise_mode) into a CUDA graph and replayed: removes the host-side launch latency that otherwise
        leaves the GPU idle between kernels.  Inputs are copied into static buffers; the returned tensors are the graph's
        static outputs and are overwritten by the next call with the same key."""
        self._on_device(ws=ws, c=c, v=v)
        R = neural_rendering_resolution or self.cfg.neural_rendering_resolution
        rk = self.rk
        key = (ws.shape[0], R, noise_mode, tuple(v.shape), rk['depth_resolution'], rk['depth_resolution_importance'], float(rk['ray_start']),
               float(rk['ray_end']), float(rk['box_warp']), bool(rk.get('white_back', False)), rk['superresolution_noise_mode'],
               bool(rk.get('sr_antialias', True)))
        g = self._graphs.get(key)
        if g is None:
            st = dict(ws=torch.empty(ws.shape, dtype=torch.float32, device=self.device), c=torch.empty(c.shape, dtype=torch.float32, device=self.device),
                      v=torch.empty(v.shape, dtype=torch.float32, device=self.device), seed=torch.zeros(1, dtype=torch.int64, device=self.device))
            st['ws'].copy_(ws); st['c'].copy_(c); st['v'].copy_(v)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                                   # warm-up outside capture (lazy attribute setup, allocator)
                self._synthesis(st['ws'], st['c'], st['v'], noise_mode, R, seed=0, seed_ptr=st['seed'])
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._synthesis(st['ws'], st['c'], st['v'], noise_mode, R, seed=0, seed_ptr=st['seed'])
            g = self._graphs[key] = (graph, st, out, self.launches, self.conv_flops)
        graph, st, out, launches, flops = g
        st['ws'].copy_(ws, non_blocking=True); st['c'].copy_(c, non_blocking=True); st['v'].copy_(v, non_blocking=True)
        st['seed'].fill_(int(seed))
        graph.replay()
        self.launches, self.conv_flops = launches, flops
        return out
