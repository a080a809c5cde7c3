/* libnext3d_b200.so -- C ABI of the Blackwell-native (sm_100a) Next3D generator hot path.
 *
 * Drop-in boundary.  The reference reaches its native code through three pybind11
 * plugins JIT-built by torch_utils/custom_ops.py:61 (`get_plugin`):
 *     bias_act_plugin.bias_act              torch_utils/ops/bias_act.cpp:36,100
 *     upfirdn2d_plugin.upfirdn2d            torch_utils/ops/upfirdn2d.cpp:20,108
 *     filtered_lrelu_plugin.filtered_lrelu  torch_utils/ops/filtered_lrelu.cpp:20,300
 * and runs everything else (modulated conv, rasterizer, grid_sample, volume renderer) through ATen /
 * cuDNN / pytorch3d / OpenCV calls.  This library replaces all of it with plain-C entry points:
 * raw device pointers + sizes in, int status out, no torch / C++ types, caller-owned buffers, explicit stream,
 * no allocation, no host sync and no global mutable device state inside (CUDA-graph capturable).
 *
 * Conventions
 *   - return value: 0 = ok, <0 = error (N3D_ERR_*); n3d_last_error() returns a thread-local message.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - "NHWC" = channels-last [N,H,W,C]; "split bf16" = two bf16 tensors (hi, lo) with hi+lo ~= the fp32 value
 *     (operands of the 3-product bf16 tensor-core scheme).
 */
#ifndef NEXT3D_B200_H
#define NEXT3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define N3D_OK 0
#define N3D_ERR_INVALID_ARG (-1)
#define N3D_ERR_UNSUPPORTED (-2)
#define N3D_ERR_CUDA (-3)

#define N3D_DTYPE_F32 0
#define N3D_DTYPE_F16 1

const char* n3d_last_error(void);
int n3d_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Op-level API: the native half of torch_utils.ops (NCHW or channels_last tensors described by element strides).
 * ---------------------------------------------------------------------------------------------------------- */

/* y = clamp(act(x + b[(i / step_b) % size_b]) * gain, +-clamp)   -- replaces bias_act_plugin.bias_act
 * (bias_act.cpp:36-94, kernel bias_act.cu:28-151; forward only, grad == 0).  act: 1 linear, 2 relu, 3 lrelu,
 * 4 tanh, 5 sigmoid, 6 elu, 7 selu, 8 softplus, 9 swish (the reference's cuda_idx, bias_act.py:23-33).
 * x dense (any memory format), b contiguous or NULL, clamp < 0 disables clamping. */
int n3d_bias_act(const void* x, const void* b, void* y, int dtype, int64_t numel, int size_b, int step_b,
                 int act, float alpha, float gain, float clamp, void* stream);

/* Zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, per (n,c) plane  -- replaces upfirdn2d_plugin.upfirdn2d
 * (upfirdn2d.cpp:20-102, kernels upfirdn2d.cu:33-204).  f: fp32 [fh,fw] (separable filters are expanded by the
 * Python wrapper).  x/y element strides in (N,C,H,W) order. */
int n3d_upfirdn2d(const void* x, const float* f, void* y, int dtype, int N, int C, int H, int W,
                  const int64_t x_strides[4], const int64_t y_strides[4], int fh, int fw, int upx, int upy,
                  int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                  int outH, int outW, void* stream);

/* bias -> FIR upsample -> lrelu*gain/clamp -> FIR downsample   -- replaces filtered_lrelu_plugin.filtered_lrelu
 * (filtered_lrelu.cpp:20-213), forward only, no sign tensor.  x [N,C,H,W] contiguous, y [N,C,outH,outW] contiguous,
 * fu [fuh,fuw] / fd [fdh,fdw] fp32 2-D filters.  tmp: caller-provided fp32 scratch of N*C*upH*upW floats where
 * upH = H*up + pady0 + pady1 - fuh + 1 (same for W). */
int n3d_filtered_lrelu(const void* x, const float* fu, const float* fd, const void* b, void* y, float* tmp, int dtype,
                       int N, int C, int H, int W, int fuh, int fuw, int fdh, int fdw, int up, int down,
                       int padx0, int padx1, int pady0, int pady1, float gain, float slope, float clamp, int flip,
                       int outH, int outW, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Generator engine: fused building blocks used by TriPlaneGenerator.synthesis (replace the ATen/cuDNN calls the
 * reference makes from networks_stylegan2.py / networks_stylegan2_styleunet.py / superresolution.py).
 * ---------------------------------------------------------------------------------------------------------- */

/* All style vectors of all modulated layers in one launch (FullyConnectedLayer affine, networks_stylegan2.py:303,315
 * and ToRGB's weight_gain :354).  Row r of the concatenated affine matrix belongs to input channel i of some layer:
 *   styles[row_ooff[r] + n * row_cin[r]] = (dot(ws[n, widx[r], :], A[r, :]) / sqrt(wdim) + b[r]) * scale[r]
 * so that every layer owns a dense [N, Cin] block (row_ooff[r] = layer_base * N + i). */
int n3d_styles(const float* ws, int N, int num_ws, int wdim, const float* affine_w, const float* affine_b,
               const int32_t* row_widx, const float* row_scale, const int64_t* row_ooff, const int32_t* row_cin,
               float* styles, int rows, void* stream);

/* Demodulation coefficients of all layers in one launch (networks_stylegan2.py:65).  Row r = output channel o of a layer:
 *   dcoef[row_ooff[r] + n*row_cout[r]] = rsqrt(sum_i styles[row_soff[r] + n*row_cin[r] + i]^2 * wsq[row_woff[r] + i] + 1e-8),
 * wsq[o,i] = sum_k W[o,i,k]^2 (dense [N, Cout] block per layer). */
int n3d_demod(const float* styles, const float* wsq, const int64_t* row_woff, const int32_t* row_cin,
              const int64_t* row_soff, const int64_t* row_ooff, const int32_t* row_cout, float* dcoef, int rows, int N,
              void* stream);

typedef struct {
    int8_t dy, dx;        /* input pixel offset of this tap relative to the output-tile pixel */
    int16_t img_off;      /* selects the input "image" group: image = n + img_off * a_img_mul (parity sub-image of stride-2 convs) */
    int32_t wtap;         /* index of the [Cout, Cin] weight slab for this tap */
} N3DConvTap;

typedef struct {
    void* hi;             /* bf16 [.., cstride] or NULL */
    void* lo;
    const float* style;   /* [N, Cout] multiplier applied before the split (next layer's modulation) or NULL */
    int32_t cstride;      /* channel stride of the destination (allows writing into a concat buffer) */
    int32_t coff;         /* channel offset in the destination */
} N3DSplitOut;

/* Optional ToRGB layer (networks_stylegan2.py:353-357) fused into the epilogue of the convolution that produces its input,
 * for image widths <= 4 (the super-resolution blocks): rgb[n,c,pixel] (+)= clamp(sum_o v[n,pixel,o] * style[n,o] * weight[c,o] + bias[c]).
 * Requires all Cout channels of a pixel in one tile (Cout <= 256) and TN == 1 tiles (resolution >= 16). out == NULL disables it. */
typedef struct {
    float* out;            /* fp32 image, NHWC [N,OH,OW,channels] or NCHW */
    const float* weight;   /* [channels, Cout] fp32 */
    const float* style;    /* [N, Cout] fp32 (ToRGB affine output, already multiplied by 1/sqrt(Cout)) */
    const float* bias;     /* [channels] */
    float clamp;           /* < 0 = none */
    int32_t channels, nchw, accumulate;
} N3DFusedRgb;

/* One implicit-GEMM convolution on tcgen05 tensor cores (bf16 hi/lo operands, 3 products, fp32 accumulate in TMEM):
 *   acc[n, y, x, o] = sum_taps sum_i A[n + tap.img_off*a_img_mul, y + tap.dy, x + tap.dx, i] * Wp[tap.wtap, o, i]
 * over an M-space of N x MH x MW "tile pixels"; out-of-range A reads are zero (this is the conv padding).
 * Epilogue (mode 0): v = acc * dcoef[n,o] + noise[oy,ox] + bias[o]; v = lrelu_slope(v) * gain; clamp; then any of
 *   out[k].hi/lo <- split_bf16(v * out[k].style[n,o]),  out_f32 <- v (optionally += existing, NHWC or NCHW).
 * Epilogue (mode 1): out_f32 <- acc (raw), used by the transposed-conv parity classes and by split-K.
 * Split-K (splits > 1): for launches with a handful of tiles and a long K loop (the 4^2 / 8^2 layers: one 128-row tile, K = 9 * 512)
 * the taps x channel-chunks are cut into `splits` slices that run as independent work items; deterministic (no atomics).
 * Output pixel of tile pixel (y, x) is (y*oy_mul + oy_off, x*ox_mul + ox_off) in an OH x OW image. */
typedef struct {
    const void* a_hi; const void* a_lo;      /* bf16 NHWC [NI, AH, AW, Cin] */
    int32_t NI, AH, AW, Cin;
    const void* w_hi; const void* w_lo;      /* bf16 [T, Cout, Cin] */
    int32_t T, Cout;
    int32_t N, MH, MW;                       /* M-space */
    int32_t a_img_mul;                       /* image coordinate = n + tap.img_off * a_img_mul (parity-major sub-images: = N) */
    int32_t ntaps; N3DConvTap taps[9];
    int32_t nprod;                           /* 3 = hi*hi + hi*lo + lo*hi (fp32-grade), 1 = hi*hi only */
    int32_t mode;
    const float* dcoef;                      /* [N, Cout] or NULL */
    const float* bias;                       /* [Cout] or NULL */
    const float* noise;                      /* [OH, OW] (already multiplied by noise_strength) or NULL */
    int64_t noise_nstride;                   /* 0 = one noise map shared by the batch ('const'), OH*OW = per-sample ('random') */
    float gain, slope, clamp;                /* slope 1 = linear; clamp < 0 = none */
    N3DSplitOut out[2];
    float* out_f32; int32_t f32_cstride, f32_coff, f32_nchw, f32_accumulate;
    int32_t oy_mul, oy_off, ox_mul, ox_off, OH, OW;
    N3DFusedRgb rgb;
    int32_t splits;                          /* > 1: split-K (mode 1 only): K slice s of every tile writes its raw partial sums to */
    int64_t split_stride;                    /*      out_f32 + s * split_stride (elements); reduce with n3d_splitk_epilogue */
} N3DConvGemm;

int n3d_conv_gemm(const N3DConvGemm* p, void* stream);

/* Stride-2 transposed 3x3 convolution (the first half of every up-sampling conv: conv_transpose2d(stride 2),
 * conv2d_resample.py:114-127) as ONE launch of the same kernel: the four output-parity classes
 *   out[2p+a, 2q+b] = sum_{ky = a (mod 2), kx = b (mod 2)} W[ky,kx] x[p - (ky-a)/2, q - (kx-b)/2]
 * are four sub-problems (4/2/2/1 taps) scheduled together.  p->MH, p->MW = INPUT height / width, p->T must be 9; the raw fp32
 * NHWC result [N, 2H+1, 2W+1, Cout] is written to p->out_f32 (f32_cstride / f32_coff honoured); taps, offsets, OH/OW, mode and
 * the epilogue fields of the descriptor are ignored (follow with n3d_fir_up_epilogue). */
int n3d_conv_transposed_gemm(const N3DConvGemm* p, void* stream);

/* fp32 NHWC -> split bf16 NHWC with optional per-(n,c) modulation: out = split(x[n,y,x,c] * style[n,c]). */
int n3d_modulate_split(const float* x, int64_t npix_per_img, int N, int C, const float* style, void* hi, void* lo,
                       int out_cstride, int out_coff, void* stream);

/* Second half of an up-sampling modulated conv (conv2d_resample.py:114-131 + networks_stylegan2.py:320-329):
 * raw [(2H+1),(2W+1)] transposed-conv output (fp32 NHWC) -> 4x4 FIR [1,3,3,1]^2/64 * 4, pad 1 -> [2H,2W], then the
 * same epilogue as n3d_conv_gemm mode 0 (requires gain > 0 and 0 <= slope <= 1: lrelu / linear). */
int n3d_fir_up_epilogue(const float* raw, int N, int H2, int W2, int C, const float* dcoef, const float* bias,
                        const float* noise, int64_t noise_nstride, float gain, float slope, float clamp,
                        const N3DSplitOut out[2], float* out_f32, int f32_cstride, int f32_coff, void* stream);

/* Second pass of a split-K convolution: y = sum_s partial[s] (fixed order) followed by the mode-0 epilogue of n3d_conv_gemm
 * (demod, noise, bias, lrelu * gain, clamp, modulated split-bf16 copies, fp32 copy).  partial: fp32 NHWC [S][N,H,W,C]. */
int n3d_splitk_epilogue(const float* partial, int S, int64_t split_stride, int N, int H, int W, int C, const float* dcoef,
                        const float* bias, const float* noise, int64_t noise_nstride, float gain, float slope, float clamp,
                        const N3DSplitOut out[2], float* out_f32, int f32_cstride, int f32_coff, void* stream);

/* First half of a down-sampling conv (conv2d_resample.py:108-111): fp32 NHWC [H,W] -> FIR pad (2,2,2,2) -> [(H+1),(W+1)]
 * -> split bf16, de-interleaved by pixel parity into 4 sub-images [4, N, SH, SW, C] (parity (y&1)*2 + (x&1), pixel
 * (y>>1, x>>1)), SH = (H+2)/2, so that the stride-2 conv becomes 9 unit-stride taps. */
int n3d_fir_down_split(const float* x, int N, int H, int W, int C, void* hi, void* lo, void* stream);

/* upfirdn2d.upsample2d on an fp32 NHWC image (up 2, pad (2,1), gain 4; networks_stylegan2.py:577); y is NHWC, or NCHW
 * when y_nchw != 0 (used for the final image). */
int n3d_upsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, int y_nchw, void* stream);
/* upfirdn2d.downsample2d on an fp32 NHWC image (pad (1,1), down 2; networks_stylegan2_styleunet.py:109). */
int n3d_downsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Mesh path: FLAME neural-texture rasterization into the orthographic views (triplane_next3d.py:190-230).
 * ---------------------------------------------------------------------------------------------------------- */

/* View transform of vertices / landmarks for the 4 rendering views (triplane_next3d.py:194-205; rotation matrices
 * are supplied by the caller, [4,3,3] fp32, computed with the reference's fp32 sin/cos):
 *   out[n, view, i, :] = flipY(p) @ R_view + shift, * scale, negate y,z; z += zoff.  Also applies the rasterizer
 * wrapper's x,y negation (renderer.py:403) when ndc_flip != 0. */
int n3d_transform_points(const float* pts, int N, int P, const float* rot, int nviews, float zoff, int ndc_flip,
                         float* out, void* stream);

/* pytorch3d-semantics rasterizer (oracle/oracle_c.c is the bit-exact CPU statement) -- replaces
 * pytorch3d.renderer.mesh.rasterize_meshes as called at volumetric_rendering/renderer.py:414-424 (blur 0, 1 face/pixel,
 * cull_backfaces, no perspective correction).  verts [NM, V, 3] NDC fp32, faces [F,3] int32 (shared by all images)
 * -> pix_to_face [NM,H,W] int32 (face id or -1), bary [NM,H,W,3] fp32 (-1 where empty).
 * workspace: caller-provided, 16-byte aligned, NM * F * 16 floats (per-face setup records shared by all pixel bins). */
int n3d_rasterize(const float* verts, const int32_t* faces, int NM, int V, int F, int H, int W, int32_t* pix_to_face,
                  float* bary, float* workspace, void* stream);

/* Texture lookup for the 4 rendered views of each sample (triplane_next3d.py:211-228, renderer.py:425-437):
 *   uv = sum_k bary_k * face_uv[f,k,:] (0 where nothing is visible); value = bilinear(texture[n], uv) (grid_sample, zeros
 *   padding, align_corners False); alpha = bilinear(eye_mask, uv) * visible.
 * pix_to_face / bary are [N*4,H,W(,3)] with image index n*4 + view; face_uv [F,3,2]; texture NHWC [N,TH,TW,C<=32];
 * eye_mask [MH,MW].  Outputs (plane-major): tex_planes [3,N,H,W,C] = (view0, view1 + view2, view3) and alpha [3,N,H,W] =
 * (view0, view1, view3) -- the reference's `alpha[1] | alpha[1]` quirk (:226) means view 2's alpha is never used. */
int n3d_uv_sample(const int32_t* pix_to_face, const float* bary, const float* face_uv, const float* texture,
                  const float* eye_mask, int N, int H, int W, int TH, int TW, int C, int MH, int MW,
                  float* tex_planes, float* alpha, void* stream);

/* fill_mouth (renderer.py:583-602): alpha [NI,H,W] in place; cv2.floodFill from (0,0) with fixed range [seed, seed+254]
 * on alpha*255, then alpha += 1 - (filled/127.5 - 1), clipped to [0,1].  H*W <= 65536.  One CTA per image. */
int n3d_fill_mouth(float* alpha, int NI, int H, int W, void* stream);

/* gen_mouth_mask (triplane_next3d.py:330-344) on device: lm2d [N,68,2] -> boxes [N,4] int32 (y0,y1,x0,x1). */
int n3d_mouth_box(const float* lm2d, int N, int32_t* boxes, void* stream);

/* Antialiased bilinear resize (ATen _upsample_bilinear2d_aa) between per-sample boxes of NHWC images --
 * replaces F.interpolate(..., antialias=True) at triplane_next3d.py:152,161 and superresolution.py:282-286:
 *   dst[n, ty0:ty1, tx0:tx1, :] = resize(src[n, sy0:sy1, sx0:sx1, :]) to (ty1-ty0, tx1-tx0).
 * src_box / dst_box: device int32 [N,4] = (y0,y1,x0,x1) or NULL (= whole image); destination pixels outside the box are
 * left untouched.  dst may be NULL when only the split output is wanted; hi/lo (optional): split_bf16(value * style[n,c]). */
int n3d_resize_aa(const float* src, int N, int SH, int SW, int C, const int32_t* src_box, float* dst, int DH, int DW,
                  const int32_t* dst_box, const float* style, void* hi, void* lo, void* stream);

/* planes[n,p,y,x,c] = tex[n,p,y,x,c] * alpha[n,p,y,x] + static[n,y,x,p*32+c] * (1 - alpha)  (triplane_next3d.py:171-174);
 * tex for plane 0 = blended_front NHWC [N,H,W,32] (neural-blending output), planes 1,2 = tex_planes[1:] (plane-major
 * [3,N,H,W,32]), alpha [3,N,H,W], static_planes NHWC [N,H,W,96]; planes out channels-last [N,3,H,W,32]. */
int n3d_blend_planes(const float* blended_front, const float* tex_planes, const float* alpha, const float* static_planes,
                     int N, int H, int W, float* planes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Volume renderer: ray generation + stratified sampling + tri-plane fetch + MLP decode + importance resampling +
 * sort-merge + alpha compositing in one kernel per ray batch (ray_sampler.py:24-63, renderer.py:95-268,
 * triplane_next3d.py:348-371, ray_marcher.py:27-66).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    const float* planes;       /* [N,3,PH,PW,32] channels-last fp32 */
    int32_t N, PH, PW;
    const float* cam2world;    /* [N,16] */
    const float* intrinsics;   /* [N,9]  */
    int32_t res;               /* rays = res*res per sample */
    int32_t depth_coarse, depth_fine;
    float ray_start, ray_end, box_warp;
    const float* u_coarse;     /* [N, M, Dc] injected uniforms or NULL (= in-kernel counter RNG with `seed`) */
    const float* u_fine;       /* [N*M, Df] or NULL */
    uint64_t seed;
    const uint64_t* seed_ptr;  /* optional device counter added to `seed` (lets a captured CUDA graph draw fresh noise per replay) */
    const float* w0; const float* b0;   /* decoder.net.0: [64,32] (already * 1/sqrt(32)), [64] */
    const float* w1; const float* b1;   /* decoder.net.2: [33,64] (already * 1/sqrt(64)), [33] */
    float* rgb;                /* [N, M, 32] */
    float* depth;              /* [N, M] (unclamped) */
    float* wsum;               /* [N, M] */
    float* depth_minmax;       /* [2] device: running min / max of all sample depths (init +inf / -inf by caller) */
    int32_t white_back;
} N3DRender;

int n3d_render_rays(const N3DRender* p, void* stream);
/* Diagnostics: micro-kernels that measure the hard per-SM limits of the renderer's formulation on the workload described by p
 * (only planes / cameras / res / depth_* / ray_* / box_warp are read; nothing is written but sink[0], which may be NULL):
 * kind 0 = the tri-plane gather alone (rays, depths, 12 taps per sample, 256-bit loads, weighted sum, bf16 split, shared-memory
 * tile) on all warps; kind 1 = the decoder's activations alone (192 MUFU operations per sample).  bench.py times them next to
 * n3d_render_rays to report how far the fused kernel is from what a per-sample gather + MLP can reach. */
int n3d_render_floor(const N3DRender* p, int kind, float* sink, void* stream);
/* composite depth clamp to the batch-global depth range (ray_marcher.py:53-54). */
int n3d_depth_clamp(float* depth, int64_t n, const float* depth_minmax, void* stream);
/* run_model only (TriPlaneGenerator.sample, triplane_next3d.py:276): coords [N,P,3] -> sigma [N,P], rgb [N,P,32] (or NULL). */
int n3d_sample_points(const float* planes, int N, int PH, int PW, const float* coords, int64_t P, float box_warp,
                      const float* w0, const float* b0, const float* w1, const float* b1, float* sigma, float* rgb,
                      void* stream);

/* Shape-extraction grid (gen_samples_next3d.py:80-102, 208-238): sigma at the voxel centres of create_samples(N = grid_n,
 * cube_length) for the flat indices head .. head+count-1, coordinates generated in the kernel with the script's own float32
 * operations (no [N^3,3] coordinate tensor), written into sigma_grid [grid_n]^3 at the position the script's
 * flip(dims=[0]) + border trim leaves them; voxels inside the trimmed border (width `pad`, the script's int(30 * N / 256))
 * receive pad_value and are not decoded.  pad = 0: plain flip.  Planes of ONE sample [3,PH,PW,32]. */
int n3d_sample_grid(const float* planes, int PH, int PW, int grid_n, float cube_length, float box_warp, int64_t head, int64_t count,
                    int pad, float pad_value, const float* w0, const float* b0, const float* w1, const float* b1, float* sigma_grid,
                    void* stream);

/* MappingNetwork + truncation in one launch (networks_stylegan2.py:233-268, called through triplane_next3d.py:111-115 with
 * c[:, :25] * c_scale): z [N,512], c [N,25] -> ws [N,num_ws,512].  embed / fc0 / fc1 are the raw parameters of
 * `backbone.mapping` (weight [out,in], bias [out]); their runtime gains (1/sqrt(in), lr_multiplier 0.01) are applied inside.
 * truncation_cutoff < 0 = all layers; w_avg may be NULL when truncation_psi == 1. */
int n3d_mapping(const float* z, const float* c, int N, float c_scale, const float* embed_w, const float* embed_b, const float* fc0_w,
                const float* fc0_b, const float* fc1_w, const float* fc1_b, const float* w_avg, float truncation_psi,
                int truncation_cutoff, int num_ws, float* ws, void* stream);

/* out[f, :] = sum_k B[f, k] * Y[k, :]  (B [F,K], Y [K,D], fp32): the frame drivers' latent interpolation (B = the cubic-spline
 * basis of gen_videos_next3d.py:106-117 evaluated once on the host) and camera smoothing (reenact_avatar_next3d.py:159) on the
 * device. */
int n3d_interp_rows(const float* B, const float* Y, int F, int K, int64_t D, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-side input parsers (no device work): what the inference scripts do per frame in Python.
 * ---------------------------------------------------------------------------------------------------------- */

/* Vertex positions of a Wavefront .obj text (gen_samples_next3d.py:165-174: every line with line[:2] == "v ", tokens after the
 * first converted with float(), all numbers flattened and reshaped to (-1, 3), then .float()).  text need not be NUL-terminated.
 * Writes up to max_vertices * 3 floats; *n_vertices receives the count (also when it exceeds max_vertices -> INVALID_ARG, so a
 * caller can size the buffer with max_vertices = 0 first). */
int n3d_parse_obj_vertices(const char* text, int64_t len, float* xyz, int64_t max_vertices, int64_t* n_vertices);

/* Whitespace-separated table of numbers with '#' comments, the subset of np.loadtxt the scripts use for *_kpt2d.txt
 * (gen_samples_next3d.py:176-177): float64 parse, float32 store; every non-empty row must have *n_cols columns. */
int n3d_parse_float_table(const char* text, int64_t len, float* values, int64_t max_values, int64_t* n_values, int64_t* n_cols);

#ifdef __cplusplus
}
#endif
#endif /* NEXT3D_B200_H */
