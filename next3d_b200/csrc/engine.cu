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

// thread = (output pixel, 4-channel group): 16 taps of float4
__global__ void __launch_bounds__(256) fir_up_epilogue_kernel(const float* __restrict__ raw, int N, int H2, int W2, int C, EpiParams E) {
    const int RH = H2 + 1, RW = W2 + 1, c4n = C >> 2;
    const int64_t total = (int64_t)N * H2 * W2 * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t t = i / c4n;
        const int x = (int)(t % W2); t /= W2;
        const int y = (int)(t % H2);
        const int n = (int)(t / H2);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int fy = 0; fy < 4; ++fy) {
            const int ry = y + fy - 1;
            if (ry < 0 || ry >= RH) continue;
#pragma unroll
            for (int fx = 0; fx < 4; ++fx) {
                const int rx = x + fx - 1;
                if (rx < 0 || rx >= RW) continue;
                const float w = (kFir1[fy] * kFir1[fx]) * 4.f;            // f/64 * gain(up^2)
                const float4 v = __ldg(reinterpret_cast<const float4*>(raw + (((int64_t)n * RH + ry) * RW + rx) * C) + c4);
                acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
            }
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        const int c0 = c4 * 4;
        const float nz = E.noise ? __ldg(E.noise + (int64_t)n * E.noise_nstride + (int64_t)y * W2 + x) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = v[j];
            if (E.dcoef) a *= __ldg(E.dcoef + (int64_t)n * C + c0 + j);
            a += nz;
            if (E.bias) a += __ldg(E.bias + c0 + j);
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
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = v[j];
                if (o.style) s *= __ldg(o.style + (int64_t)n * C + c0 + j);
                split_bf16(s, h[j], l[j]);
            }
            const int64_t off = pix * o.cstride + o.coff + c0;
            *reinterpret_cast<uint2*>((__nv_bfloat16*)o.hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
            *reinterpret_cast<uint2*>((__nv_bfloat16*)o.lo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
        }
    }
}

// FIR (pad 2,2,2,2) -> [(H+1),(W+1)] -> parity-split bf16 hi/lo, layout [4 parities][N][SH][SW][C]
__global__ void __launch_bounds__(256) fir_down_split_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int FH = H + 1, FW = W + 1, SH = (H + 2) / 2, SW = (W + 2) / 2, c4n = C >> 2;
    const int64_t total = (int64_t)4 * N * SH * SW * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t t = i / c4n;
        const int sx = (int)(t % SW); t /= SW;
        const int sy = (int)(t % SH); t /= SH;
        const int n = (int)(t % N);
        const int par = (int)(t / N);
        const int y = sy * 2 + (par >> 1), xx = sx * 2 + (par & 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < FH && xx < FW) {
#pragma unroll
            for (int fy = 0; fy < 4; ++fy) {
                const int iy = y + fy - 2;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int fx = 0; fx < 4; ++fx) {
                    const int ix = xx + fx - 2;
                    if (ix < 0 || ix >= W) continue;
                    const float w = kFir1[fy] * kFir1[fx];
                    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * H + iy) * W + ix) * C) + c4);
                    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
                }
            }
        }
        __nv_bfloat16 h[4], l[4];
        split_bf16(acc.x, h[0], l[0]); split_bf16(acc.y, h[1], l[1]); split_bf16(acc.z, h[2], l[2]); split_bf16(acc.w, h[3], l[3]);
        const int64_t o = i * 4;   // same linearisation as the output layout
        *reinterpret_cast<uint2*>(hi + o) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(lo + o) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
}

// upsample2d: zero-insert x2, pad (2,1), 4x4 FIR * 4  ->  out[y,x] = sum over input taps with matching parity
__global__ void __launch_bounds__(256) upsample2d_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ y, int nchw) {
    const int OH = 2 * H, OW = 2 * W;
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
            const int uy = oy + fy - 2;                    // position in the zero-inserted image
            if (uy < 0 || (uy & 1) || (uy >> 1) >= H) continue;
#pragma unroll
            for (int fx = 0; fx < 4; ++fx) {
                const int ux = ox + fx - 2;
                if (ux < 0 || (ux & 1) || (ux >> 1) >= W) continue;
                acc += (kFir1[fy] * kFir1[fx]) * 4.f * __ldg(x + (((int64_t)n * H + (uy >> 1)) * W + (ux >> 1)) * C + c);
            }
        }
        if (nchw) y[(((int64_t)n * C + c) * OH + oy) * OW + ox] = acc;
        else y[i] = acc;
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
    const int64_t total = (int64_t)N * H2 * W2 * (C / 4);
    fir_up_epilogue_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(raw, N, H2, W2, C, E);
    N3D_CHECK_LAUNCH("n3d_fir_up_epilogue");
    return N3D_OK;
}

extern "C" int n3d_fir_down_split(const float* x, int N, int H, int W, int C, void* hi, void* lo, void* stream) {
    N3D_CHECK_ARG(x && hi && lo, "n3d_fir_down_split: null pointer");
    N3D_CHECK_ARG(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "n3d_fir_down_split: C %% 4 and even H, W required");
    const int64_t total = (int64_t)4 * N * ((H + 2) / 2) * ((W + 2) / 2) * (C / 4);
    fir_down_split_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    N3D_CHECK_LAUNCH("n3d_fir_down_split");
    return N3D_OK;
}

extern "C" int n3d_upsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, int y_nchw, void* stream) {
    N3D_CHECK_ARG(x && y, "n3d_upsample2d_nhwc: null pointer");
    upsample2d_kernel<<<grid_for((int64_t)N * 4 * H * W * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y, y_nchw);
    N3D_CHECK_LAUNCH("n3d_upsample2d_nhwc");
    return N3D_OK;
}

extern "C" int n3d_downsample2d_nhwc(const float* x, int N, int H, int W, int C, float* y, void* stream) {
    N3D_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0, "n3d_downsample2d_nhwc: bad args");
    downsample2d_kernel<<<grid_for((int64_t)N * (H / 2) * (W / 2) * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, y);
    N3D_CHECK_LAUNCH("n3d_downsample2d_nhwc");
    return N3D_OK;
}
