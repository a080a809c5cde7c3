// C entry points of the volume renderer and of point decoding (kernels: render_fused.cu) plus the depth clamp of the ray
// marcher (ray_marcher.py:53-54).
#include "common.cuh"
#include "../../include/next3d_b200.h"
#include <stdlib.h>

namespace {
constexpr int kMaxD = 96;       // depth samples per pass the fused kernel lays out (3 tiles x 32 lanes per ray)

__global__ void __launch_bounds__(256) depth_clamp_kernel(float* __restrict__ depth, int64_t n, const float* __restrict__ mm) {
    const float lo = mm[0], hi = mm[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float d = depth[i];
        if (isnan(d)) d = INFINITY;
        depth[i] = fminf(fmaxf(d, lo), hi);
    }
}
}  // namespace

int n3d_render_fused_launch(const N3DRender* p, void* stream, int mode);     // render_fused.cu
int n3d_render_floor_launch(const N3DRender* p, int kind, float* sink, void* stream);
int n3d_points_fused_launch(const float* planes, int N, int PH, int PW, const float* coords, long long Pn, float box_warp, const float* w0,
                            const float* b0, const float* w1, const float* b1, float* sigma, float* rgb, int grid_n, float cube_length, long long head,
                            int pad, float pad_value, void* stream);

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
    N3D_CHECK_ARG(planes && coords && w0 && b0 && w1 && b1 && sigma && N >= 1 && P >= 1 && box_warp > 0.f, "n3d_sample_points: bad args");
    N3D_CHECK_ARG((long long)N * 3 * PH * PW * 128 < (1ll << 32), "n3d_sample_points: plane tensor too large for 32-bit texel offsets");
    return n3d_points_fused_launch(planes, N, PH, PW, coords, P, box_warp, w0, b0, w1, b1, sigma, rgb, 0, 0.f, 0, 0, 0.f, stream);
}

extern "C" int n3d_sample_grid(const float* planes, int PH, int PW, int grid_n, float cube_length, float box_warp, int64_t head, int64_t count,
                               int pad, float pad_value, const float* w0, const float* b0, const float* w1, const float* b1, float* sigma_grid,
                               void* stream) {
    N3D_CHECK_ARG(planes && w0 && b0 && w1 && b1 && sigma_grid && grid_n >= 2 && cube_length > 0.f && box_warp > 0.f, "n3d_sample_grid: bad args");
    N3D_CHECK_ARG(head >= 0 && count >= 1 && head + count <= (int64_t)grid_n * grid_n * grid_n, "n3d_sample_grid: index range outside the grid");
    N3D_CHECK_ARG(pad >= 0 && 2 * pad <= grid_n, "n3d_sample_grid: bad trim width");
    N3D_CHECK_ARG((long long)3 * PH * PW * 128 < (1ll << 32), "n3d_sample_grid: plane tensor too large for 32-bit texel offsets");
    return n3d_points_fused_launch(planes, 1, PH, PW, nullptr, count, box_warp, w0, b0, w1, b1, sigma_grid, nullptr, grid_n, cube_length, head, pad,
                                   pad_value, stream);
}

extern "C" int n3d_render_floor(const N3DRender* p, int kind, float* sink, void* stream) {
    N3D_CHECK_ARG(p && p->planes && p->cam2world && p->intrinsics && (kind == 0 || kind == 1), "n3d_render_floor: bad args");
    N3D_CHECK_ARG(p->depth_coarse >= 4 && p->depth_coarse <= kMaxD && p->depth_fine >= 0 && p->depth_fine <= kMaxD, "n3d_render_floor: depth resolutions");
    return n3d_render_floor_launch(p, kind, sink, stream);
}
