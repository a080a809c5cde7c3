// Memory-bound glue kernels of the generator engine (everything between two tensor-core convolutions that could not be
// folded into a GEMM epilogue).  All tensors NHWC; 128-bit accesses along the channel dimension; grids sized in
// multiples of the SM count.
#include "common.cuh"
#include "../../include/next3d_b200.h"

namespace {
constexpr int kSMs = 148;
inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
    int64_t g = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)kSMs * per_sm;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------- styles / demod
// one warp per (row, n): 512-long dot product with float4 loads
// output layout: every layer owns a dense [N, Cin] block: styles[ooff[r] + n * cin[r]]
__global__ void __launch_bounds__(256) styles_kernel(const float* __restrict__ ws, int N, int num_ws, int wdim,
                                                     const float* __restrict__ A, const float* __restrict__ b,
                                                     const int* __restrict__ widx, const float* __restrict__ scale,
                                                     const int64_t* __restrict__ ooff, const int* __restrict__ cin,
                                                     float* __restrict__ styles, int rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv = rsqrtf((float)wdim);
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        const float4* a4 = reinterpret_cast<const float4*>(A + r * wdim);
        const int wi = widx[r];
        for (int n = 0; n < N; ++n) {
            const float4* w4 = reinterpret_cast<const float4*>(ws + ((int64_t)n * num_ws + wi) * wdim);
            float acc = 0.f;
            for (int i = lane; i < wdim / 4; i += 32) {
                const float4 a = __ldg(a4 + i), w = __ldg(w4 + i);
                acc += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
            }
            acc = warp_sum(acc);
            if (lane == 0) styles[ooff[r] + (int64_t)n * cin[r]] = (acc * inv + b[r]) * scale[r];
        }
    }
}

// one warp per output channel row r of some layer: styles of that layer at styles[soff[r] + n*cin[r] + i],
// result at dcoef[ooff[r] + n*cout[r]] (dense [N, Cout] block per layer)
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ styles, const float* __restrict__ wsq,
                                                    const int64_t* __restrict__ woff, const int* __restrict__ cin,
                                                    const int64_t* __restrict__ soff, const int64_t* __restrict__ ooff,
                                                    const int* __restrict__ cout, float* __restrict__ dcoef, int rows, int N) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        const float* w = wsq + woff[r];
        const int c = cin[r];
        for (int n = 0; n < N; ++n) {
            const float* s = styles + soff[r] + (int64_t)n * c;
            float acc = 0.f;
            for (int i = lane; i < c; i += 32) { const float sv = s[i]; acc += sv * sv * __ldg(w + i); }
            acc = warp_sum(acc);
            if (lane == 0) dcoef[ooff[r] + (int64_t)n * cout[r]] = rsqrtf(acc + 1e-8f);
        }
    }
}

// ---------------------------------------------------------------------------------------------- modulate + split
__global__ void __launch_bounds__(256) modulate_split_kernel(const float* __restrict__ x, int64_t npix, int N, int C,
                                                             const float* __restrict__ style, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, int cstride, int coff) {
    const int c4n = C >> 2;
    const int64_t total = (int64_t)N * npix * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const int64_t pix = i / c4n;
        const int n = (int)(pix / npix);
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        if (style) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(style + (int64_t)n * C) + c4);
            v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
        }
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        const int64_t o = pix * cstride + coff + c4 * 4;
        *reinterpret_cast<uint2*>(hi + o) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(lo + o) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
}

// ---------------------------------------------------------------------------------------------- FIR helpers
// [1,3,3,1]/8 per axis; upfirdn2d's 2-D filter is the outer product / 64 (setup_filter, upfirdn2d.py:101-111)
__device__ __constant__ float kFir1[4] = {0.125f, 0.375f, 0.375f, 0.125f};   // immutable constants (not runtime state)

struct EpiParams {
    const float* dcoef; const float* bias; const float* noise; int64_t noise_nstride;
    float gain, slope, clamp;
    N3DSplitOut out[2];
    float* out_f32; int f32_cstride, f32_coff;
};

__device__ __forceinline__ float4 f4_fma(float w, const float4& v, const float4& a) { return make_float4(fmaf(w, v.x, a.x), fmaf(w, v.y, a.y), fmaf(w, v.z, a.z), fmaf(w, v.w, a.w)); }
__device__ __forceinline__ float4 f4_scale(float w, const float4& v) { return make_float4(w * v.x, w * v.y, w * v.z, w * v.w); }

__device__ __forceinline__ void epi_store(const EpiParams& E, float4 acc, int n, int y, int x, int H2, int W2, int C, int c0) {
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    const float nz = E.noise ? __ldg(E.noise + (int64_t)n * E.noise_nstride + (int64_t)y * W2 + x) : 0.f;
    float4 dc = make_float4(1.f, 1.f, 1.f, 1.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (E.dcoef) dc = __ldg(reinterpret_cast<const float4*>(E.dcoef + (int64_t)n * C + c0));
    if (E.bias) bs = __ldg(reinterpret_cast<const float4*>(E.bias + c0));
    const float dcv[4] = {dc.x, dc.y, dc.z, dc.w}, bsv[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[j] * dcv[j] + nz + bsv[j];
        a = (a > 0.f ? a : a * E.slope) * E.gain;
        if (E.clamp >= 0.f) a = fminf(fmaxf(a, -E.clamp), E.clamp);
        v[j] = a;
    }
    const int64_t pix = ((int64_t)n * H2 + y) * W2 + x;
    if (E.out_f32) *reinterpret_cast<float4*>(E.out_f32 + pix * E.f32_cstride + E.f32_coff + c0) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const N3DSplitOut o = E.out[k];
        if (!o.hi) continue;
        float4 st = make_float4(1.f, 1.f, 1.f, 1.f);
        if (o.style) st = __ldg(reinterpret_cast<const float4*>(o.style + (int64_t)n * C + c0));
        const float sv[4] = {st.x, st.y, st.z, st.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(v[j] * sv[j], h[j], l[j]);
        const int64_t off = pix * o.cstride + o.coff + c0;
        *reinterpret_cast<uint2*>((__nv_bfloat16*)o.hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>((__nv_bfloat16*)o.lo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
}

// thread = (2x2 block of output pixels, 4-channel group).  The 4x4 FIR is separable ([1,3,3,1]/8 * 2 per axis): five raw rows
// of five float4 each are streamed through registers, each row contributing to the two output rows -> 25 loads per 4 outputs
// instead of 64 (the straightforward form is L1-bandwidth bound).
__global__ void __launch_bounds__(256) fir_up_epilogue_kernel(const float* __restrict__ raw, int N, int H2, int W2, int C, EpiParams E) {
    const int RH = H2 + 1, RW = W2 + 1, c4n = C >> 2, BH = H2 >> 1, BW = W2 >> 1;
    const int64_t total = (int64_t)N * BH * BW * c4n;
    const float g[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t t = i / c4n;
        const int bx = (int)(t % BW); t /= BW;
        const int by = (int)(t % BH);
        const int n = (int)(t / BH);
        const int y0 = by * 2, x0 = bx * 2;
        float4 acc[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) acc[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int ry = y0 - 1 + r;
            if (ry < 0 || ry >= RH) continue;
            float4 v[5];
#pragma unroll
            for (int cidx = 0; cidx < 5; ++cidx) {
                const int rx = x0 - 1 + cidx;
                v[cidx] = (rx >= 0 && rx < RW) ? __ldg(reinterpret_cast<const float4*>(raw + (((int64_t)n * RH + ry) * RW + rx) * C) + c4)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 h0 = f4_scale(g[0], v[0]), h1 = f4_scale(g[0], v[1]);
#pragma unroll
            for (int fx = 1; fx < 4; ++fx) { h0 = f4_fma(g[fx], v[fx], h0); h1 = f4_fma(g[fx], v[fx + 1], h1); }
            if (r < 4) { acc[0][0] = f4_fma(g[r], h0, acc[0][0]); acc[0][1] = f4_fma(g[r], h1, acc[0][1]); }
            if (r > 0) { acc[1][0] = f4_fma(g[r - 1], h0, acc[1][0]); acc[1][1] = f4_fma(g[r - 1], h1, acc[1][1]); }
        }
        const int c0 = c4 * 4;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) epi_store(E, acc[a][b], n, y0 + a, x0 + b, H2, W2, C, c0);
    }
}

// FIR (pad 2,2,2,2) -> [(H+1),(W+1)] -> parity-split bf16 hi/lo, layout [4 parities][N][SH][SW][C].
// thread = (sub-pixel (sy,sx), 4-channel group) -> the 2x2 block of FIR outputs (2sy+a, 2sx+b), one per parity image; separable
// row streaming as above (25 loads per 4 outputs).
__global__ void __launch_bounds__(256) fir_down_split_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int SH = (H + 2) / 2, SW = (W + 2) / 2, c4n = C >> 2;
    const int64_t total = (int64_t)N * SH * SW * c4n;
    const int64_t par_stride = (int64_t)N * SH * SW * C;
    const float g[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t t = i / c4n;
        const int sx = (int)(t % SW); t /= SW;
        const int sy = (int)(t % SH);
        const int n = (int)(t / SH);
        const int y0 = sy * 2, x0 = sx * 2;                  // FIR output (y, x) reads input rows y-2 .. y+1
        float4 acc[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) acc[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int iy = y0 - 2 + r;
            if (iy < 0 || iy >= H) continue;
            float4 v[5];
#pragma unroll
            for (int cidx = 0; cidx < 5; ++cidx) {
                const int ix = x0 - 2 + cidx;
                v[cidx] = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * H + iy) * W + ix) * C) + c4)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 h0 = f4_scale(g[0], v[0]), h1 = f4_scale(g[0], v[1]);
#pragma unroll
            for (int fx = 1; fx < 4; ++fx) { h0 = f4_fma(g[fx], v[fx], h0); h1 = f4_fma(g[fx], v[fx + 1], h1); }
            if (r < 4) { acc[0][0] = f4_fma(g[r], h0, acc[0][0]); acc[0][1] = f4_fma(g[r], h1, acc[0][1]); }
            if (r > 0) { acc[1][0] = f4_fma(g[r - 1], h0, acc[1][0]); acc[1][1] = f4_fma(g[r - 1], h1, acc[1][1]); }
        }
        const int64_t o = (((int64_t)n * SH + sy) * SW + sx) * C + c4 * 4;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float4 v = acc[a][b];
                if (y0 + a > H || x0 + b > W) v = make_float4(0.f, 0.f, 0.f, 0.f);     // beyond the (H+1)x(W+1) FIR output: zero pad
                __nv_bfloat16 h[4], l[4];
                split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
                const int64_t off = (int64_t)(a * 2 + b) * par_stride + o;
                *reinterpret_cast<uint2*>(hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
                *reinterpret_cast<uint2*>(lo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
            }
    }
}

// upsample2d: zero-insert x2, pad (2,1), 4x4 FIR * 4.  Polyphase form: out[2i] = .25 x[i-1] + .75 x[i], out[2i+1] = .75 x[i] + .25 x[i+1]
// per axis.  thread = (input pixel (i,j), VEC channels) -> the 2x2 output block from the 3x3 input neighbourhood.
template <int VEC>
__global__ void __launch_bounds__(256) upsample2d_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ y, int nchw) {
    const int OH = 2 * H, OW = 2 * W, cvn = C / VEC;
    const int64_t total = (int64_t)N * H * W * cvn;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % cvn);
        int64_t t = idx / cvn;
        const int j = (int)(t % W); t /= W;
        const int i = (int)(t % H);
        const int n = (int)(t / H);
        float v[3][3][VEC];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int yy = i - 1 + a, xx = j - 1 + b;
                const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
                const float* src = x + (((int64_t)n * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * C + cv * VEC;
                if (VEC == 4) {
                    const float4 q = ok ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[a][b][0] = q.x; v[a][b][1 % VEC] = q.y; v[a][b][2 % VEC] = q.z; v[a][b][3 % VEC] = q.w;
                } else {
                    v[a][b][0] = ok ? __ldg(src) : 0.f;
                }
            }
        const float wa[2][3] = {{0.25f, 0.75f, 0.f}, {0.f, 0.75f, 0.25f}};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float o[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    float acc = 0.f;
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) acc += (wa[a][p] * wa[b][q]) * v[p][q][e];
                    o[e] = acc;
                }
                const int oy = 2 * i + a, ox = 2 * j + b;
                if (nchw) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) y[(((int64_t)n * C + cv * VEC + e) * OH + oy) * OW + ox] = o[e];
                } else {
                    float* dst = y + (((int64_t)n * OH + oy) * OW + ox) * C + cv * VEC;
                    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
                    else dst[0] = o[0];
                }
            }
    }
}

// downsample2d: pad (1,1), 4x4 FIR, keep every 2nd pixel
__global__ void __launch_bounds__(256) downsample2d_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ y) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % OW); t /= OW;
        const int oy = (int)(t % OH);
        const int n = (int)(t / OH);
        float acc = 0.f;
#pragma unroll
        for (int fy = 0; fy < 4; ++fy) {
            const int iy = oy * 2 + fy - 1;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int fx = 0; fx < 4; ++fx) {
                const int ix = ox * 2 + fx - 1;
                if (ix < 0 || ix >= W) continue;
                acc += (kFir1[fy] * kFir1[fx]) * __ldg(x + (((int64_t)n * H + iy) * W + ix) * C + c);
            }
        }
        y[i] = acc;
    }
}
}  // namespace

extern "C" int n3d_styles(const float* ws, int N, int num_ws, int wdim, const float* affine_w, const float* affine_b,
                          const int32_t* row_widx, const float* row_scale, const int64_t* row_ooff, const int32_t* row_cin,
                          float* styles, int rows, void* stream) {
    N3D_CHECK_ARG(ws && affine_w && affine_b && row_widx && row_scale && row_ooff && row_cin && styles, "n3d_styles: null pointer");
    N3D_CHECK_ARG(wdim % 4 == 0 && rows > 0 && N > 0, "n3d_styles: bad sizes");
    styles_kernel<<<grid_for((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(ws, N, num_ws, wdim, affine_w, affine_b, row_widx,
                                                                                      row_scale, row_ooff, row_cin, styles, rows);
    N3D_CHECK_LAUNCH("n3d_styles");
    return N3D_OK;
}

extern "C" int n3d_demod(const float* styles, const float* wsq, const int64_t* row_woff, const int32_t* row_cin,
                         const int64_t* row_soff, const int64_t* row_ooff, const int32_t* row_cout, float* dcoef, int rows, int N,
                         void* stream) {
    N3D_CHECK_ARG(styles && wsq && row_woff && row_cin && row_soff && row_ooff && row_cout && dcoef, "n3d_demod: null pointer");
    demod_kernel<<<grid_for((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(styles, wsq, row_woff, row_cin, row_soff, row_ooff,
                                                                                     row_cout, dcoef, rows, N);
    N3D_CHECK_LAUNCH("n3d_demod");
    return N3D_OK;
}

extern "C" int n3d_modulate_split(const float* x, int64_t npix_per_img, int N, int C, const float* style, void* hi, void* lo,
                                  int out_cstride, int out_coff, void* stream) {
    N3D_CHECK_ARG(x && hi && lo, "n3d_modulate_split: null pointer");
    N3D_CHECK_ARG(C % 4 == 0 && out_cstride % 4 == 0 && out_coff % 4 == 0, "n3d_modulate_split: channels must be multiples of 4");
    const int64_t total = (int64_t)N * npix_per_img * (C / 4);
    modulate_split_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, npix_per_img, N, C, style, (__nv_bfloat16*)hi,
                                                                                     (__nv_bfloat16*)lo, out_cstride, out_coff);
    N3D_CHECK_LAUNCH("n3d_modulate_split");
    return N3D_OK;
}

extern "C" int n3d_fir_up_epilogue(const float* raw, int N, int H2, int W2, int C, const float* dcoef, const float* bias,
                                   const float* noise, int64_t noise_nstride, float gain, float slope, float clamp,
                                   const N3DSplitOut out[2], float* out_f32, int f32_cstride, int f32_coff, void* stream) {
    N3D_CHECK_ARG(raw && out, "n3d_fir_up_epilogue: null pointer");
    N3D_CHECK_ARG(C % 4 == 0, "n3d_fir_up_epilogue: C must be a multiple of 4");
    EpiParams E;
    E.dcoef = dcoef; E.bias = bias; E.noise = noise; E.noise_nstride = noise_nstride; E.gain = gain; E.slope = slope; E.clamp = clamp;
    E.out[0] = out[0]; E.out[1] = out[1]; E.out_f32 = out_f32; E.f32_cstride = f32_cstride; E.f32_coff = f32_coff;
    for (int k = 0; k < 2; ++k)
        N3D_CHECK_ARG(!E.out[k].hi || ((E.out[k].cstride % 4 == 0) && (E.out[k].coff % 4 == 0)), "n3d_fir_up_epilogue: unaligned split output");
    N3D_CHECK_ARG(!out_f32 || (f32_cstride % 4 == 0 && f32_coff % 4 == 0), "n3d_fir_up_epilogue: unaligned fp32 output");
    N3D_CHECK_ARG(H2 % 2 == 0 && W2 % 2 == 0, "n3d_fir_up_epilogue: output size must be even");
    const int64_t total = (int64_t)N * (H2 / 2) * (W2 / 2) * (C / 4);
    fir_up_epilogue_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(raw, N, H2, W2, C, E);
    N3D_CHECK_LAUNCH("n3d_fir_up_epilogue");
    return N3D_OK;
}

extern "C" int n3d_fir_down_split(const float* x, int N, int H, int W, int C, void* hi, void* lo, void* stream) {
    N3D_CHECK_ARG(x && hi && lo, "n3d_fir_down_split: null pointer");
    N3D_CHECK_ARG(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "n3d_fir_down_split: C %% 4 and even H, W required");
    const int64_t total = (int64_t)N * ((H + 2) / 2) * ((W + 2) / 2) * (C / 4);
    fir_down_split_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    N3D_CHECK_LAUNCH("n3d_fir_down_split");
    return N3D_OK;
}

extern "C" int n3d_upsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, int y_nchw, void* stream) {
    N3D_CHECK_ARG(x && y, "n3d_upsample2d_nhwc: null pointer");
    if (C % 4 == 0)
        upsample2d_kernel<4><<<grid_for((int64_t)N * H * W * (C / 4), 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, y_nchw);
    else
        upsample2d_kernel<1><<<grid_for((int64_t)N * H * W * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, y_nchw);
    N3D_CHECK_LAUNCH("n3d_upsample2d_nhwc");
    return N3D_OK;
}

extern "C" int n3d_downsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, void* stream) {
    N3D_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0, "n3d_downsample2d_nhwc: bad args");
    downsample2d_kernel<<<grid_for((int64_t)N * (H / 2) * (W / 2) * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y);
    N3D_CHECK_LAUNCH("n3d_downsample2d_nhwc");
    return N3D_OK;
}
