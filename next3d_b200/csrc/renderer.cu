// Fused volume renderer: ray generation, stratified depths, tri-plane bilinear fetch (+ mean over planes), MLP decoder
// (32 -> 64 softplus -> 33, sigmoid), coarse compositing weights, importance resampling, sort-merge of coarse + fine
// samples and final alpha compositing -- one kernel, nothing but the final [N,M,32] / depth / weight-sum leaves the SM.
// Replaces RaySampler.forward (ray_sampler.py:24-63), ImportanceRenderer.forward (renderer.py:95-268), OSGDecoder.forward
// (triplane_next3d.py:359-371) and MipRayMarcher2.run_forward (ray_marcher.py:27-66), which materialise
// [N,3,M*D,32] feature tensors (604 MB per pass at batch 8) in the reference.
//
// Organisation (see the comment above render_kernel; run_model on arbitrary points keeps the SIMT decoder below):
//   CTA = RAYS rays x D samples = up to 192 threads.  Gather: one warp per sample, lanes = the 32 channels, so each of the
//   12 bilinear taps is one coalesced 128-byte line of the channels-last planes.  Decode: one thread per sample, weights
//   broadcast from shared memory as float4.  Per-ray work (compositing weights via a warp product scan, CDF build + inversion,
//   rank-counting sort-merge, colour accumulation) runs one warp per ray.
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

// ---- warp-cooperative per-ray primitives (one warp owns one ray; lanes stride over the samples) ----
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MipRayMarcher2 weights (ray_marcher.py:27-46) of a depth-sorted sample list held in shared memory:
//   alpha_k = 1 - exp(-softplus((s_k + s_k+1)/2 - 1) * (t_k+1 - t_k)),  T_k = prod_{i<k} (1 - alpha_i + 1e-10),  w_k = alpha_k T_k
// computed by the whole warp: alphas in parallel, the transmittance by a chunked warp product scan.  Returns sum_k w_k.
__device__ __forceinline__ float warp_march_weights(int cnt, const float* __restrict__ t, const float* __restrict__ sg, float* __restrict__ w, int lane) {
    float carry = 1.f, wsum = 0.f;
    for (int base = 0; base < cnt - 1; base += 32) {
        const int k = base + lane;
        float alpha = 0.f;
        if (k < cnt - 1) {
            const float delta = t[k + 1] - t[k];
            const float dens = softplus_t((sg[k] + sg[k + 1]) / 2.f - 1.f);
            alpha = 1.f - expf(-(dens * delta));
        }
        float f = k < cnt - 1 ? (1.f - alpha + 1e-10f) : 1.f;
        float inc = f;                                   // inclusive product scan over the 32 lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc *= up;
        }
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.f;
        const float T = carry * exc;
        if (k < cnt - 1) { const float wk = alpha * T; w[k] = wk; wsum += wk; }
        carry *= __shfl_sync(0xffffffffu, inc, 31);
    }
    return warp_sum_f(wsum);
}

// ---------------------------------------------------------------------------------------------------------------------
// render_kernel: one CTA = R rays x D samples (<= 128 decoded samples per pass = one UMMA M tile), 256 threads.
//   gather   : warp per sample, lanes = channels; the tri-plane feature is written straight into the K-major SWIZZLE_64B
//              bf16 (hi, lo) A-operand tile of layer 1;
//   decoder  : tcgen05.mma, bf16x3 (hi*hi + hi*lo + lo*hi), fp32 accumulators in TMEM:
//                layer 1  [128 x 32] x [32 x 64]  -> TMEM cols 0..63   (6 MMAs, N = 64)
//                epilogue 1: + b0, softplus, split -> A tile of layer 2 (SWIZZLE_128B) written by the row's thread
//                layer 2  [128 x 64] x [64 x 48]  -> TMEM cols 64..111 (12 MMAs, N = 48, 33 used)
//                epilogue 2: + b1, sigma / sigmoid colours -> fp32 rows [33] in shared memory;
//   per ray  : one warp per ray (weights by warp scans, inverse-CDF sampling, rank-counting sort-merge, compositing).
// Operand tiles are written with ordinary shared-memory stores using the same XOR swizzle TMA would apply
// (16-byte chunk index ^ row bits), then fence.proxy.async + barrier before the single MMA-issuing thread runs.
constexpr int kRThreads = 256;
constexpr int kTileRows = 128;
constexpr int kN2 = 48;                                   // layer-2 UMMA N (33 outputs padded to a multiple of 16)

__device__ __forceinline__ uint32_t sw64_off(int row, int byte_in_row) {       // 64-byte rows, Swizzle<2,4,3>
    return (uint32_t)(row * 64 + ((((byte_in_row >> 4) ^ ((row >> 1) & 3)) << 4) | (byte_in_row & 15)));
}
__device__ __forceinline__ uint32_t sw128_off(int row, int byte_in_row) {      // 128-byte rows, Swizzle<3,4,3>
    return (uint32_t)(row * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15)));
}

__global__ void __launch_bounds__(kRThreads, 2) render_kernel(const RenderK K) {
    using namespace n3d_tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const N3DRender& P = K.p;
    const int Dc = P.depth_coarse, Df = P.depth_fine, R = K.rays_per_cta;
    const int Dt = Dc + Df;
    // ---- shared memory carve-up: operand tiles first (1024-byte aligned), then fp32 scratch
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* opF_hi = base;                               // [128][32] bf16, 64-B rows   (8 KiB)
    uint8_t* opF_lo = opF_hi + kTileRows * 64;
    uint8_t* opH_hi = opF_lo + kTileRows * 64;            // [128][64] bf16, 128-B rows  (16 KiB)
    uint8_t* opH_lo = opH_hi + kTileRows * 128;
    uint8_t* opW0_hi = opH_lo + kTileRows * 128;          // [64][32] bf16 (4 KiB)
    uint8_t* opW0_lo = opW0_hi + kHidden * 64;
    uint8_t* opW1_hi = opW0_lo + kHidden * 64;            // [48][64] bf16 (6 KiB)
    uint8_t* opW1_lo = opW1_hi + kN2 * 128;
    float* sB0 = reinterpret_cast<float*>(opW1_lo + kN2 * 128);   // [64]
    float* sB1 = sB0 + kHidden;                          // [48]
    float* sC = sB1 + kN2;                               // coarse rows [R*Dc][33]
    float* sF = sC + R * Dc * kRowStride;                // fine rows   [R*Df][33]
    float* sTc = sF + R * Df * kRowStride;               // coarse depths [R][Dc]
    float* sTf = sTc + R * Dc;                           // fine depths   [R][Df]
    float* sWgt = sTf + R * Df;                          // weights [R][Dt]
    float* sScr = sWgt + R * Dt;                         // per-ray scratch [R][3*Dt]: sorted depths | sorted sigmas | cdf
    float* sRay = sScr + R * 3 * Dt;                     // [R][8]: origin xyz, dir xyz, image index
    unsigned char* sOrd = reinterpret_cast<unsigned char*>(sRay + R * 8);   // [R][Dt] merged order
    __shared__ float s_min[kRThreads / 32], s_max[kRThreads / 32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint64_t seed = P.seed_ptr ? (P.seed + __ldg(reinterpret_cast<const unsigned long long*>(P.seed_ptr))) : P.seed;
    const int64_t total_rays = (int64_t)P.N * K.M;
    const int64_t ngroups = (total_rays + R - 1) / R;
    const uint32_t bar = smem_u32(&s_bar);

    // ---- one-time setup: mbarrier, TMEM (128 columns), decoder weights -> bf16 (hi, lo) B-operand tiles
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < kHidden * kFeat; i += blockDim.x) {             // W0 [64][32] (row = hidden unit n, K = 32)
        const int n = i / kFeat, k = i - n * kFeat;
        __nv_bfloat16 h, l;
        split_bf16(__ldg(P.w0 + i), h, l);
        *reinterpret_cast<__nv_bfloat16*>(opW0_hi + sw64_off(n, k * 2)) = h;
        *reinterpret_cast<__nv_bfloat16*>(opW0_lo + sw64_off(n, k * 2)) = l;
    }
    for (int i = tid; i < kN2 * kHidden; i += blockDim.x) {               // W1 [48][64] (rows >= 33 are zero)
        const int n = i / kHidden, k = i - n * kHidden;
        __nv_bfloat16 h, l;
        split_bf16(n < kOut ? __ldg(P.w1 + n * kHidden + k) : 0.f, h, l);
        *reinterpret_cast<__nv_bfloat16*>(opW1_hi + sw128_off(n, k * 2)) = h;
        *reinterpret_cast<__nv_bfloat16*>(opW1_lo + sw128_off(n, k * 2)) = l;
    }
    for (int i = tid; i < kHidden; i += blockDim.x) sB0[i] = __ldg(P.b0 + i);
    for (int i = tid; i < kN2; i += blockDim.x) sB1[i] = i < kOut ? __ldg(P.b1 + i) : 0.f;
    // rows of the A tiles that no sample maps to are still multiplied: keep them finite
    for (int i = tid; i < (kTileRows * 64) / 16; i += blockDim.x) {
        reinterpret_cast<uint4*>(opF_hi)[i] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(opF_lo)[i] = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < (kTileRows * 128) / 16; i += blockDim.x) {
        reinterpret_cast<uint4*>(opH_hi)[i] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(opH_lo)[i] = make_uint4(0, 0, 0, 0);
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const float scale = 2.f / P.box_warp;
    const int64_t plane_img = (int64_t)3 * P.PH * P.PW * kFeat;
    const uint64_t dhi64 = umma_desc_hi(32), dhi128 = umma_desc_hi(64);
    uint32_t mma_phase = 0;
    float dmin = INFINITY, dmax = -INFINITY;

    // persistent CTA: decoder operands, TMEM and the mbarrier are set up once, then ray groups are processed in a loop
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
    const int64_t ray0 = group * R;                       // global ray index = n*M + m
    // ---- rays (ray_sampler.py:43-63)
    if (tid < R) {
        const int64_t gr = ray0 + tid;
        if (gr < total_rays) {
            const int n = (int)(gr / K.M), m = (int)(gr % K.M);
            const int i = m / P.res, j = m % P.res;
            const float inv = 1.f / (float)P.res, half = 0.5f / (float)P.res;
            const float xc = (float)j * inv + half, yc = (float)i * inv + half;
            const float* I = P.intrinsics + n * 9;
            const float fx = I[0], sk = I[1], cx = I[2], fy = I[4], cy = I[5];
            const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx;
            const float yl = (yc - cy) / fy;
            const float* C = P.cam2world + n * 16;
            float wv[3], o[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                wv[a] = C[a * 4] * xl + C[a * 4 + 1] * yl + C[a * 4 + 2] + C[a * 4 + 3];
                o[a] = C[a * 4 + 3];
                wv[a] -= o[a];
            }
            const float nrm = fmaxf(sqrtf(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]), 1e-12f);
            float* r = sRay + tid * 8;
            r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
            r[3] = wv[0] / nrm; r[4] = wv[1] / nrm; r[5] = wv[2] / nrm;
        } else {
            float* r = sRay + tid * 8;
            r[0] = r[1] = r[2] = r[3] = r[4] = r[5] = 0.f;
        }
        // sample index -> image of this ray (tail rays of the last CTA reuse the last valid ray; their results are never stored)
        sRay[tid * 8 + 6] = __int_as_float((int)(min(gr, total_rays - 1) / K.M));
    }
    // ---- coarse depths (renderer.py:203-205)
    for (int s = tid; s < R * Dc; s += blockDim.x) {
        const int r = s / Dc, k = s % Dc;
        const int64_t gr = ray0 + r;
        float t = 0.f;
        if (gr < total_rays) {
            const float u = P.u_coarse ? __ldg(P.u_coarse + gr * Dc + k) : hash_uniform(seed, (uint64_t)(gr * Dc + k));
            t = linspace_at(P.ray_start, P.ray_end, Dc, k) + u * K.delta_coarse;
        }
        sTc[s] = t;
    }
    __syncthreads();

    // gather + decode of `cnt` samples (depths in sT, R rays x D samples) into rows[cnt][33]
    auto gather_decode = [&](const float* sT, int D, float* rows) {
        const int cnt = R * D;
        {
            int r = warp / D, k = warp - r * D;                  // sample s = r * D + k, advanced without divisions
#pragma unroll 2
            for (int s = warp; s < cnt; s += nwarps) {
                const float* ry = sRay + r * 8;
                const float t = sT[s];
                const float feat = triplane_feature(P.planes + (int64_t)__float_as_int(ry[6]) * plane_img, P.PH, P.PW, ry[0] + t * ry[3],
                                                    ry[1] + t * ry[4], ry[2] + t * ry[5], scale, lane);
                __nv_bfloat16 h, l;
                split_bf16(feat, h, l);
                const uint32_t off = sw64_off(s, lane * 2);
                *reinterpret_cast<__nv_bfloat16*>(opF_hi + off) = h;
                *reinterpret_cast<__nv_bfloat16*>(opF_lo + off) = l;
                k += nwarps;
                while (k >= D) { k -= D; ++r; }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        // ---- layer 1 on the tensor cores
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(kHidden);
            const uint32_t a_hi = smem_u32(opF_hi), a_lo = smem_u32(opF_lo), b_hi = smem_u32(opW0_hi), b_lo = smem_u32(opW0_lo);
#pragma unroll
            for (int k16 = 0; k16 < kFeat / 16; ++k16) {
                const uint32_t ko = (uint32_t)k16 * 32u;
                umma_bf16(tmem, umma_desc(a_hi + ko, dhi64), umma_desc(b_hi + ko, dhi64), idesc, k16 != 0);
                umma_bf16(tmem, umma_desc(a_hi + ko, dhi64), umma_desc(b_lo + ko, dhi64), idesc, 1u);
                umma_bf16(tmem, umma_desc(a_lo + ko, dhi64), umma_desc(b_hi + ko, dhi64), idesc, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, mma_phase, nullptr, 0);
        mma_phase ^= 1u;
        tc_fence_after();
        // ---- epilogue 1: thread = row (TMEM lane); softplus(acc + b0) -> split -> layer-2 A tile
        if (tid < kTileRows) {
            const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c16 = 0; c16 < kHidden; c16 += 16) {
                uint32_t rr[16];
                tmem_ld16(t_row + (uint32_t)c16, rr);
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(softplus_fast(__uint_as_float(rr[j]) + sB0[c16 + j]), h0, l0);
                    split_bf16(softplus_fast(__uint_as_float(rr[j + 1]) + sB0[c16 + j + 1]), h1, l1);
                    ph[j >> 1] = pack_bf16x2(h0, h1);
                    pl[j >> 1] = pack_bf16x2(l0, l1);
                }
                if (tid < cnt) {
                    const uint32_t o0 = sw128_off(tid, c16 * 2), o1 = sw128_off(tid, c16 * 2 + 16);
                    *reinterpret_cast<uint4*>(opH_hi + o0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4*>(opH_hi + o1) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    *reinterpret_cast<uint4*>(opH_lo + o0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    *reinterpret_cast<uint4*>(opH_lo + o1) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                }
            }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        // ---- layer 2
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(kN2);
            const uint32_t a_hi = smem_u32(opH_hi), a_lo = smem_u32(opH_lo), b_hi = smem_u32(opW1_hi), b_lo = smem_u32(opW1_lo);
            const uint32_t d = tmem + (uint32_t)kHidden;
#pragma unroll
            for (int k16 = 0; k16 < kHidden / 16; ++k16) {
                const uint32_t ko = (uint32_t)k16 * 32u;
                umma_bf16(d, umma_desc(a_hi + ko, dhi128), umma_desc(b_hi + ko, dhi128), idesc, k16 != 0);
                umma_bf16(d, umma_desc(a_hi + ko, dhi128), umma_desc(b_lo + ko, dhi128), idesc, 1u);
                umma_bf16(d, umma_desc(a_lo + ko, dhi128), umma_desc(b_hi + ko, dhi128), idesc, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, mma_phase, nullptr, 0);
        mma_phase ^= 1u;
        tc_fence_after();
        // ---- epilogue 2: sigma = o[0], rgb = sigmoid(o[1..32]) * 1.002 - 0.001 (triplane_next3d.py:369-370)
        if (tid < kTileRows) {
            const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)kHidden;
            float* row = rows + tid * kRowStride;
#pragma unroll
            for (int c16 = 0; c16 < kN2; c16 += 16) {
                uint32_t rr[16];
                tmem_ld16(t_row + (uint32_t)c16, rr);
                if (tid < cnt) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int o = c16 + j;
                        if (o >= kOut) continue;
                        const float val = __uint_as_float(rr[j]) + sB1[o];
                        if (o == 0) row[32] = val;
                        else row[o - 1] = sigmoid_fast(val) * 1.002f - 0.001f;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    };

    // ---- coarse pass
    gather_decode(sTc, Dc, sC);

    // ---- per ray (one warp each): coarse weights -> smoothed pdf -> inverse-CDF samples (renderer.py:209-268)
    for (int r = warp; r < R; r += nwarps) {
        const int64_t gr = ray0 + r;
        if (gr >= total_rays || Df <= 0) continue;
        float* w = sWgt + r * Dt;                // weights [Dc-1]
        float* sg = sScr + r * 3 * Dt;           // coarse sigmas gathered contiguously
        float* cdf = sg + 2 * Dt;                // [nw+1]
        const float* tc = sTc + r * Dc;
        const float* rows = sC + (r * Dc) * kRowStride;
        for (int k = lane; k < Dc; k += 32) sg[k] = rows[k * kRowStride + 32];
        __syncwarp();
        warp_march_weights(Dc, tc, sg, w, lane);
        __syncwarp();
        // max_pool1d(k2,s1,pad1) -> avg_pool1d(k2,s1) -> +0.01 ; keep entries [1:-1] => Dc-3 pdf weights, + 1e-5
        const int nw = Dc - 3;
        float part = 0.f;
        for (int i = lane; i < nw; i += 32) {
            const int q = i + 1;
            const float mp0 = fmaxf(w[q - 1], w[q]);
            const float mp1 = q + 1 <= Dc - 2 ? fmaxf(w[q], w[q + 1]) : w[q];
            const float v = (mp0 + mp1) * 0.5f + 0.01f + 1e-5f;
            cdf[i + 1] = v;
            part += v;
        }
        const float total = warp_sum_f(part);
        __syncwarp();
        float carry = 0.f;                       // inclusive sum scan of pdf = v / total
        for (int base = 0; base < nw; base += 32) {
            const int i = base + lane;
            float inc = i < nw ? cdf[i + 1] / total : 0.f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            if (i < nw) cdf[i + 1] = carry + inc;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cdf[0] = 0.f;
        __syncwarp();
        float* tf = sTf + r * Df;
        for (int j = lane; j < Df; j += 32) {
            const float u = P.u_fine ? __ldg(P.u_fine + gr * Df + j) : hash_uniform(seed ^ 0xA5A5A5A5DEADBEEFull, (uint64_t)(gr * Df + j));
            int lo = 0, hi = nw + 1;             // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
            const int below = max(lo - 1, 0), above = min(lo, nw);
            const float cb = cdf[below], ca = cdf[above];
            const float bb = 0.5f * (tc[below] + tc[below + 1]), ba = 0.5f * (tc[above] + tc[above + 1]);
            float denom = ca - cb;
            if (denom < 1e-5f) denom = 1.f;
            tf[j] = bb + (u - cb) / denom * (ba - bb);
        }
    }
    __syncthreads();

    // ---- fine pass
    gather_decode(sTf, Df, sF);

    // ---- per ray (one warp each): stable sort-merge of coarse (already sorted) and fine depths by rank counting
    //      (torch.sort on the concatenation, renderer.py:164-182: ties keep coarse before fine and fine in input order),
    //      final weights, composite depth and colours (ray_marcher.py:27-66)
    for (int r = warp; r < R; r += nwarps) {
        const int64_t gr = ray0 + r;
        if (gr >= total_rays) continue;
        const float* tc = sTc + r * Dc;
        const float* tf = sTf + r * Df;
        const float* rc = sC + (r * Dc) * kRowStride;
        const float* rf = sF + (r * Df) * kRowStride;
        unsigned char* ord = sOrd + r * Dt;
        float* sd = sScr + r * 3 * Dt;           // sorted depths
        float* sg = sd + Dt;                     // sorted sigmas
        float* w = sWgt + r * Dt;
        for (int a = lane; a < Dc; a += 32) {    // coarse sample a lands after every strictly smaller fine sample
            const float v = tc[a];
            int pos = a;
            for (int i = 0; i < Df; ++i) pos += tf[i] < v ? 1 : 0;
            ord[pos] = (unsigned char)a; sd[pos] = v; sg[pos] = rc[a * kRowStride + 32];
        }
        for (int j = lane; j < Df; j += 32) {    // fine sample j: rank among the fine ones (stable) + coarse samples <= it
            const float v = tf[j];
            int pos = 0;
            for (int i = 0; i < Df; ++i) pos += (tf[i] < v || (tf[i] == v && i < j)) ? 1 : 0;
            for (int a = 0; a < Dc; ++a) pos += tc[a] <= v ? 1 : 0;
            ord[pos] = (unsigned char)(Dc + j); sd[pos] = v; sg[pos] = rf[j * kRowStride + 32];
        }
        __syncwarp();
        const float wsum = warp_march_weights(Dt, sd, sg, w, lane);
        __syncwarp();
        float dacc = 0.f;
        for (int k = lane; k < Dt - 1; k += 32) dacc += w[k] * ((sd[k] + sd[k + 1]) / 2.f);
        dacc = warp_sum_f(dacc);
        if (lane == 0) {
            float depth = dacc / wsum;
            if (isnan(depth)) depth = INFINITY;                     // nan_to_num(nan=inf); the clamp kernel finishes the job
            P.depth[gr] = depth;
            P.wsum[gr] = wsum;
            dmin = fminf(sd[0], dmin);
            dmax = fmaxf(sd[Dt - 1], dmax);
        }
        // colours: lanes = channels
        auto color_at = [&](int k) { const int o = ord[k]; return o < Dc ? rc[o * kRowStride + lane] : rf[(o - Dc) * kRowStride + lane]; };
        float acc = 0.f;
        float c_prev = color_at(0);
        for (int k = 0; k < Dt - 1; ++k) {
            const float c_next = color_at(k + 1);
            acc += w[k] * ((c_prev + c_next) / 2.f);
            c_prev = c_next;
        }
        if (P.white_back) acc = acc + 1.f - wsum;
        P.rgb[gr * kFeat + lane] = acc * 2.f - 1.f;
    }
    __syncthreads();

    }   // ray-group loop

    // ---- batch-global depth range (ray_marcher.py:54): warp + block reduce, then one atomic pair per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0) { s_min[warp] = dmin; s_max[warp] = dmax; }
    __syncthreads();
    if (tid == 0 && P.depth_minmax) {
        for (int i = 1; i < nwarps; ++i) { dmin = fminf(dmin, s_min[i]); dmax = fmaxf(dmax, s_max[i]); }
        // depths are positive (ray_start > 0): IEEE ordering == signed-int ordering
        if (dmin < INFINITY) atomicMin(reinterpret_cast<int*>(P.depth_minmax), __float_as_int(dmin));
        if (dmax > -INFINITY) atomicMax(reinterpret_cast<int*>(P.depth_minmax) + 1, __float_as_int(dmax));
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
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

size_t render_smem_bytes(int R, int Dc, int Df) {
    const int Dt = Dc + Df;
    const size_t operands = (size_t)2 * kTileRows * 64 + 2 * kTileRows * 128 + 2 * kHidden * 64 + 2 * kN2 * 128;      // F, H, W0, W1 (hi + lo)
    const size_t fl = (size_t)kHidden + kN2 + (size_t)R * Dc * kRowStride + (size_t)R * Df * kRowStride + (size_t)R * Dc + (size_t)R * Df +
                      (size_t)R * Dt * 4 + (size_t)R * 8;
    return 1024 + operands + fl * sizeof(float) + (size_t)R * Dt + 16;
}
}  // namespace

extern "C" int n3d_render_rays(const N3DRender* p, void* stream) {
    N3D_CHECK_ARG(p && p->planes && p->cam2world && p->intrinsics && p->w0 && p->b0 && p->w1 && p->b1 && p->rgb && p->depth && p->wsum,
                  "n3d_render_rays: null pointer");
    N3D_CHECK_ARG(p->depth_coarse >= 4 && p->depth_coarse <= kMaxD && p->depth_fine >= 0 && p->depth_fine <= kMaxD,
                  "n3d_render_rays: depth resolutions (%d, %d) outside [4, %d]", p->depth_coarse, p->depth_fine, kMaxD);
    N3D_CHECK_ARG(p->depth_fine == 0 || p->depth_fine >= p->depth_coarse - 2, "n3d_render_rays: depth_fine must be >= depth_coarse - 2 (scratch layout)");
    N3D_CHECK_ARG(p->depth_fine > 0, "n3d_render_rays: depth_fine == 0 (coarse-only rendering) is not supported yet");
    N3D_CHECK_ARG(p->res >= 1 && p->N >= 1 && p->ray_start > 0.f && p->ray_end > p->ray_start, "n3d_render_rays: bad ray setup");
    RenderK K;
    K.p = *p;
    K.M = p->res * p->res;
    const int dmax = p->depth_coarse > p->depth_fine ? p->depth_coarse : p->depth_fine;
    K.rays_per_cta = kTileRows / dmax;                          // all samples of a pass form one 128-row UMMA tile
    if (K.rays_per_cta < 1) K.rays_per_cta = 1;
    K.delta_coarse = (float)(((double)p->ray_end - (double)p->ray_start) / (double)(p->depth_coarse - 1));
    const int threads = kRThreads;
    const size_t smem = render_smem_bytes(K.rays_per_cta, p->depth_coarse, p->depth_fine);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            n3d_set_error("n3d_render_rays: cannot raise dynamic shared memory");
            return N3D_ERR_CUDA;
        }
        configured = true;
    }
    const int64_t total_rays = (int64_t)p->N * K.M;
    const int64_t ngroups = (total_rays + K.rays_per_cta - 1) / K.rays_per_cta;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const int grid = (int)(ngroups < 2 * num_sms ? ngroups : 2 * num_sms);       // persistent: 2 CTAs per SM loop over the ray groups
    render_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(K);
    N3D_CHECK_LAUNCH("n3d_render_rays");
    return N3D_OK;
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
