// Fused volume renderer: ray generation, stratified depths, tri-plane bilinear fetch (+ mean over planes), MLP decoder
// (32 -> 64 softplus -> 33, sigmoid), coarse compositing weights, importance resampling, sort-merge of coarse + fine
// samples and final alpha compositing -- one kernel, nothing but the final [N,M,32] / depth / weight-sum leaves the SM.
// Replaces RaySampler.forward (ray_sampler.py:24-63), ImportanceRenderer.forward (renderer.py:95-268), OSGDecoder.forward
// (triplane_next3d.py:359-371) and MipRayMarcher2.run_forward (ray_marcher.py:27-66), which materialise
// [N,3,M*D,32] feature tensors (604 MB per pass at batch 8) in the reference.
//
// v1 organisation (SIMT fp32, exact-math transcendental functions):
//   CTA = RAYS rays x D samples = up to 192 threads.  Gather: one warp per sample, lanes = the 32 channels, so each of the
//   12 bilinear taps is one coalesced 128-byte line of the channels-last planes.  Decode: one thread per sample, weights
//   broadcast from shared memory as float4.  Per-ray work (compositing weights via a warp product scan, CDF build + inversion,
//   rank-counting sort-merge, colour accumulation) runs one warp per ray.
#include "common.cuh"
#include "../../include/next3d_b200.h"
#include <stdlib.h>

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

__global__ void __launch_bounds__(kMaxThreads) render_kernel(const RenderK K) {
    extern __shared__ __align__(16) float smem[];
    const N3DRender& P = K.p;
    const int Dc = P.depth_coarse, Df = P.depth_fine, R = K.rays_per_cta;
    const int Dt = Dc + Df;
    // ---- shared memory carve-up
    float* sW0 = smem;                                   // [64][32]
    float* sW1t = sW0 + kHidden * kFeat;                 // [64][36]
    float* sB0 = sW1t + kHidden * kW1Stride;             // [64]
    float* sB1 = sB0 + kHidden;                          // [36]
    float* sC = sB1 + kW1Stride;                         // coarse rows [R*Dc][33]
    float* sF = sC + R * Dc * kRowStride;                // fine rows   [R*Df][33]
    float* sTc = sF + R * Df * kRowStride;               // coarse depths [R][Dc]
    float* sTf = sTc + R * Dc;                           // fine depths   [R][Df]
    float* sWgt = sTf + R * Df;                          // weights [R][Dt]
    float* sScr = sWgt + R * Dt;                         // per-ray scratch [R][3*Dt]: sorted depths | sorted sigmas | cdf
    float* sRay = sScr + R * 3 * Dt;                     // [R][8]: origin xyz, dir xyz
    unsigned char* sOrd = reinterpret_cast<unsigned char*>(sRay + R * 8);   // [R][Dt] merged order
    __shared__ float s_min[kMaxThreads / 32], s_max[kMaxThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint64_t seed = P.seed_ptr ? (P.seed + __ldg(reinterpret_cast<const unsigned long long*>(P.seed_ptr))) : P.seed;
    const int64_t ray0 = (int64_t)blockIdx.x * R;        // global ray index = n*M + m
    const int64_t total_rays = (int64_t)P.N * K.M;

    load_decoder(P.w0, P.b0, P.w1, P.b1, sW0, sB0, sW1t, sB1);

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

    const float scale = 2.f / P.box_warp;
    const int64_t plane_img = (int64_t)3 * P.PH * P.PW * kFeat;

    // ---- coarse pass: gather (warp per sample) then decode (thread per sample)
    {
        int r = warp / Dc, k = warp - r * Dc;                  // sample s = r * Dc + k, advanced without divisions
#pragma unroll 2
        for (int s = warp; s < R * Dc; s += nwarps) {
            const float* ry = sRay + r * 8;
            const float t = sTc[s];
            sC[s * kRowStride + lane] = triplane_feature(P.planes + (int64_t)__float_as_int(ry[6]) * plane_img, P.PH, P.PW, ry[0] + t * ry[3],
                                                         ry[1] + t * ry[4], ry[2] + t * ry[5], scale, lane);
            k += nwarps;
            while (k >= Dc) { k -= Dc; ++r; }
        }
    }
    __syncthreads();
    for (int s = tid; s < R * Dc; s += blockDim.x) decode_row(sC + s * kRowStride, sW0, sB0, sW1t, sB1);
    __syncthreads();

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
    {
        int r = warp / Df, k = warp - r * Df;
#pragma unroll 2
        for (int s = warp; s < R * Df; s += nwarps) {
            const float* ry = sRay + r * 8;
            const float t = sTf[s];
            sF[s * kRowStride + lane] = triplane_feature(P.planes + (int64_t)__float_as_int(ry[6]) * plane_img, P.PH, P.PW, ry[0] + t * ry[3],
                                                         ry[1] + t * ry[4], ry[2] + t * ry[5], scale, lane);
            k += nwarps;
            while (k >= Df) { k -= Df; ++r; }
        }
    }
    __syncthreads();
    for (int s = tid; s < R * Df; s += blockDim.x) decode_row(sF + s * kRowStride, sW0, sB0, sW1t, sB1);
    __syncthreads();

    // ---- per ray (one warp each): stable sort-merge of coarse (already sorted) and fine depths by rank counting
    //      (torch.sort on the concatenation, renderer.py:164-182: ties keep coarse before fine and fine in input order),
    //      final weights, composite depth and colours (ray_marcher.py:27-66)
    float dmin = INFINITY, dmax = -INFINITY;
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
    size_t fl = (size_t)kHidden * kFeat + kHidden * kW1Stride + kHidden + kW1Stride + (size_t)R * Dc * kRowStride + (size_t)R * Df * kRowStride +
                (size_t)R * Dc + (size_t)R * Df + (size_t)R * Dt * 4 + (size_t)R * 8;
    return fl * sizeof(float) + (size_t)R * Dt + 16;
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
    static int target_threads = 0;
    if (!target_threads) {
        const char* e = getenv("N3D_RENDER_THREADS");          // tuning knob: threads (= rays x samples) per CTA
        target_threads = e ? atoi(e) : kMaxThreads;
        if (target_threads < 32 || target_threads > kMaxThreads) target_threads = kMaxThreads;
    }
    K.rays_per_cta = target_threads / dmax;
    if (K.rays_per_cta < 1) K.rays_per_cta = 1;
    K.delta_coarse = (float)(((double)p->ray_end - (double)p->ray_start) / (double)(p->depth_coarse - 1));
    const int threads = ((K.rays_per_cta * dmax + 31) / 32) * 32;
    const size_t smem = render_smem_bytes(K.rays_per_cta, p->depth_coarse, p->depth_fine);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            n3d_set_error("n3d_render_rays: cannot raise dynamic shared memory");
            return N3D_ERR_CUDA;
        }
        configured = 200 * 1024;
    }
    const int64_t total_rays = (int64_t)p->N * K.M;
    const int grid = (int)((total_rays + K.rays_per_cta - 1) / K.rays_per_cta);
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
