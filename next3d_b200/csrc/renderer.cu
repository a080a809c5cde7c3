// Point decoding on arbitrary coordinates (run_model, renderer.py:149-155 / triplane_next3d.py:232-276), the depth clamp of the
// ray marcher (ray_marcher.py:53-54) and the C entry point of the volume renderer, whose kernel lives in render_fused.cu.
#include "common.cuh"
#include "../../include/next3d_b200.h"
#include <stdlib.h>
#include "tc_ptx.cuh"

namespace {

constexpr int kMaxThreads = 192;
constexpr int kFeat = 32;
constexpr int kHidden = 64;
constexpr int kOut = 33;
constexpr int kRowStride = 33;          // 32 colours + sigma; odd stride => conflict-free per-thread rows
constexpr int kW1Stride = 36;           // transposed layer-2 weights [64][36] (33 used), float4-aligned
constexpr int kMaxD = 96;

struct RenderK {
    N3DRender p;
    int rays_per_cta;
    int M;
    float delta_coarse;
};

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch Softplus(beta 1, threshold 20)
// MUFU-based variants for the 96 activations per decoded sample (ex2.approx / lg2.approx / rcp.approx): absolute error ~1e-7 on
// O(1) values, far inside the 2e-5 kernel tolerance; the few per-ray compositing transcendentals keep the exact versions.
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

// torch.linspace(start, end, steps) for float32: symmetric evaluation around the midpoint
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
    const float step = (end - start) / (float)(steps - 1);
    return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

// Tri-plane feature of one sample, computed by a whole warp (lanes = the 32 channels).
// The 12 bilinear taps (3 planes x 4 corners; plane 0 <- (x,y), 1 <- (x,z), 2 <- (z,y); grid_sample with zeros padding,
// align_corners=False) are set up ONCE per sample by lanes 0..11 -- lane l owns plane l/4, corner l%4 and computes that tap's
// element offset (clamped, always valid) and weight (0 when out of range) -- and then broadcast with shuffles, instead of every
// lane redundantly running the whole address/weight arithmetic.  All 12 loads (each one coalesced 128-byte line) are issued
// before any is consumed.  Returns ((f0 + f1) + f2) / 3 like sampled_features.mean(1).
__device__ __forceinline__ float triplane_feature(const float* __restrict__ planes_n, int PH, int PW, float px, float py, float pz, float scale, int lane) {
    const int plane = (lane >> 2) % 3, corner = lane & 3;
    const float x = scale * px, y = scale * py, z = scale * pz;
    const float gx = plane == 2 ? z : x;
    const float gy = plane == 1 ? z : y;
    const float ix = ((gx + 1.f) * (float)PW - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)PH - 1.f) * 0.5f;
    const float flx = floorf(ix), fly = floorf(iy);
    const int xi = (int)flx + (corner & 1), yi = (int)fly + (corner >> 1);
    const float wx = (corner & 1) ? ix - flx : 1.f - (ix - flx);
    const float wy = (corner >> 1) ? iy - fly : 1.f - (iy - fly);
    const bool inside = xi >= 0 && xi < PW && yi >= 0 && yi < PH;
    const float my_w = inside ? wx * wy : 0.f;
    const int my_off = ((plane * PH + min(max(yi, 0), PH - 1)) * PW + min(max(xi, 0), PW - 1)) * kFeat;
    float v[12], w[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int off = __shfl_sync(0xffffffffu, my_off, i);
        w[i] = __shfl_sync(0xffffffffu, my_w, i);
        v[i] = __ldg(planes_n + (unsigned)(off + lane));
    }
    float f[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) a += w[p * 4 + c] * v[p * 4 + c];
        f[p] = a;
    }
    return ((f[0] + f[1]) + f[2]) / 3.f;
}

// decode one sample in place: row[0..31] features -> row[0..31] rgb, row[32] sigma
__device__ __forceinline__ void decode_row(float* __restrict__ row, const float* __restrict__ sW0, const float* __restrict__ sB0,
                                           const float* __restrict__ sW1t, const float* __restrict__ sB1) {
    float f[kFeat];
#pragma unroll
    for (int c = 0; c < kFeat; ++c) f[c] = row[c];
    float o[kW1Stride];
#pragma unroll
    for (int j = 0; j < kW1Stride; ++j) o[j] = j < kOut ? sB1[j] : 0.f;
#pragma unroll 2
    for (int j = 0; j < kHidden; ++j) {
        float a = sB0[j];
        const float4* w = reinterpret_cast<const float4*>(sW0 + j * kFeat);
#pragma unroll
        for (int c4 = 0; c4 < kFeat / 4; ++c4) {
            const float4 wv = w[c4];
            a = fmaf(wv.x, f[c4 * 4], a); a = fmaf(wv.y, f[c4 * 4 + 1], a); a = fmaf(wv.z, f[c4 * 4 + 2], a); a = fmaf(wv.w, f[c4 * 4 + 3], a);
        }
        const float h = softplus_fast(a);
        const float4* w1 = reinterpret_cast<const float4*>(sW1t + j * kW1Stride);
#pragma unroll
        for (int o4 = 0; o4 < kW1Stride / 4; ++o4) {
            const float4 wv = w1[o4];
            o[o4 * 4] += wv.x * h; o[o4 * 4 + 1] += wv.y * h; o[o4 * 4 + 2] += wv.z * h; o[o4 * 4 + 3] += wv.w * h;
        }
    }
#pragma unroll
    for (int c = 0; c < kFeat; ++c) row[c] = sigmoid_fast(o[1 + c]) * 1.002f - 0.001f;
    row[32] = o[0];
}

__device__ __forceinline__ void load_decoder(const float* w0, const float* b0, const float* w1, const float* b1, float* sW0, float* sB0,
                                             float* sW1t, float* sB1) {
    for (int i = threadIdx.x; i < kHidden * kFeat; i += blockDim.x) sW0[i] = w0[i];
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) sB0[i] = b0[i];
    for (int i = threadIdx.x; i < kHidden * kW1Stride; i += blockDim.x) {
        const int j = i / kW1Stride, o = i % kW1Stride;
        sW1t[i] = o < kOut ? w1[o * kHidden + j] : 0.f;
    }
    for (int i = threadIdx.x; i < kW1Stride; i += blockDim.x) sB1[i] = i < kOut ? b1[i] : 0.f;
}

__global__ void __launch_bounds__(256) depth_clamp_kernel(float* __restrict__ depth, int64_t n, const float* __restrict__ mm) {
    const float lo = mm[0], hi = mm[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float d = depth[i];
        if (isnan(d)) d = INFINITY;
        depth[i] = fminf(fmaxf(d, lo), hi);
    }
}

// run_model on arbitrary points: CTA of 192 threads handles 192 points
__global__ void __launch_bounds__(kMaxThreads) sample_points_kernel(const float* __restrict__ planes, int N, int PH, int PW, const float* __restrict__ coords,
                                                                   int64_t Pn, float scale, const float* w0, const float* b0, const float* w1,
                                                                   const float* b1, float* __restrict__ sigma, float* __restrict__ rgb) {
    extern __shared__ __align__(16) float smem[];
    float* sW0 = smem;
    float* sW1t = sW0 + kHidden * kFeat;
    float* sB0 = sW1t + kHidden * kW1Stride;
    float* sB1 = sB0 + kHidden;
    float* sC = sB1 + kW1Stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    load_decoder(w0, b0, w1, b1, sW0, sB0, sW1t, sB1);
    __syncthreads();
    const int64_t total = (int64_t)N * Pn;
    const int64_t plane_img = (int64_t)3 * PH * PW * kFeat;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
        for (int s = warp; s < (int)blockDim.x; s += nwarps) {
            const int64_t g = base + s;
            float feat = 0.f;
            if (g < total) {
                const float* c = coords + g * 3;
                feat = triplane_feature(planes + (g / Pn) * plane_img, PH, PW, __ldg(c), __ldg(c + 1), __ldg(c + 2), scale, lane);
            }
            sC[s * kRowStride + lane] = feat;
        }
        __syncthreads();
        if (base + tid < total) {
            decode_row(sC + tid * kRowStride, sW0, sB0, sW1t, sB1);
            sigma[base + tid] = sC[tid * kRowStride + 32];
        }
        __syncthreads();
        if (rgb) {
            for (int s = warp; s < (int)blockDim.x; s += nwarps)
                if (base + s < total) rgb[(base + s) * kFeat + lane] = sC[s * kRowStride + lane];
        }
        __syncthreads();
    }
}

}  // namespace

int n3d_render_fused_launch(const N3DRender* p, void* stream, int mode);     // render_fused.cu

extern "C" int n3d_render_rays(const N3DRender* p, void* stream) {
    N3D_CHECK_ARG(p && p->planes && p->cam2world && p->intrinsics && p->w0 && p->b0 && p->w1 && p->b1 && p->rgb && p->depth && p->wsum,
                  "n3d_render_rays: null pointer");
    if (p->depth_coarse > kMaxD || p->depth_fine > kMaxD) {
        n3d_set_error("n3d_render_rays: depth resolutions (%d, %d) above %d samples per pass are not supported", p->depth_coarse, p->depth_fine, kMaxD);
        return N3D_ERR_UNSUPPORTED;
    }
    N3D_CHECK_ARG(p->depth_coarse >= 4 && p->depth_fine >= 0, "n3d_render_rays: depth resolutions (%d, %d): need coarse >= 4, fine >= 0",
                  p->depth_coarse, p->depth_fine);
    N3D_CHECK_ARG(p->res >= 1 && p->N >= 1 && p->ray_start > 0.f && p->ray_end > p->ray_start && p->box_warp > 0.f, "n3d_render_rays: bad ray setup");
    N3D_CHECK_ARG((long long)p->N * 3 * p->PH * p->PW * 128 < (1ll << 32), "n3d_render_rays: plane tensor too large for 32-bit texel offsets");
    static const int mode = getenv("N3D_RENDER_MODE") ? atoi(getenv("N3D_RENDER_MODE")) : 0;      // diagnostics only (phase floors)
    return n3d_render_fused_launch(p, stream, mode);
}

extern "C" int n3d_depth_clamp(float* depth, int64_t n, const float* depth_minmax, void* stream) {
    N3D_CHECK_ARG(depth && depth_minmax && n >= 0, "n3d_depth_clamp: bad args");
    if (n == 0) return N3D_OK;
    depth_clamp_kernel<<<(int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth, n, depth_minmax);
    N3D_CHECK_LAUNCH("n3d_depth_clamp");
    return N3D_OK;
}

extern "C" int n3d_sample_points(const float* planes, int N, int PH, int PW, const float* coords, int64_t P, float box_warp,
                                 const float* w0, const float* b0, const float* w1, const float* b1, float* sigma, float* rgb,
                                 void* stream) {
    N3D_CHECK_ARG(planes && coords && w0 && b0 && w1 && b1 && sigma && N >= 1 && P >= 1, "n3d_sample_points: bad args");
    const size_t smem = ((size_t)kHidden * kFeat + kHidden * kW1Stride + kHidden + kW1Stride + (size_t)kMaxThreads * kRowStride) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(sample_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
            n3d_set_error("n3d_sample_points: cannot raise dynamic shared memory");
            return N3D_ERR_CUDA;
        }
        configured = true;
    }
    const int64_t total = (int64_t)N * P;
    int64_t grid = (total + kMaxThreads - 1) / kMaxThreads;
    if (grid > 148 * 16) grid = 148 * 16;
    sample_points_kernel<<<(int)grid, kMaxThreads, smem, (cudaStream_t)stream>>>(planes, N, PH, PW, coords, P, 2.f / box_warp, w0, b0, w1, b1, sigma, rgb);
    N3D_CHECK_LAUNCH("n3d_sample_points");
    return N3D_OK;
}
