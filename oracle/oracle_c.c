/* CPU oracle, integer/graph parts in plain C.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * n3d_oracle_rasterize : restatement of pytorch3d `rasterize_meshes` (naive per-pixel algorithm) for the
 *   settings the reference uses at volumetric_rendering/renderer.py:389-397,414-424 (blur_radius 0,
 *   faces_per_pixel 1, perspective_correct False, cull_backfaces True, clip_barycentric_coords False).
 *   pytorch3d is NOT vendored in /root/reference and is pinned nowhere (environment.yml:15-38), so this
 *   follows its published algorithm as written down in SURVEY.md Appendix C -- "parity unpinned".
 *   All arithmetic is IEEE fp32 in the expression order written here; build with -ffp-contract=off so
 *   that the CUDA kernel (which uses __fmul_rn/__fsub_rn/...) can match it bit for bit.
 * n3d_oracle_floodfill : cv2.floodFill(seed (0,0), newVal 255, loDiff 0, upDiff 254,
 *   FLOODFILL_FIXED_RANGE, 4-connectivity) as called by fill_mouth, renderer.py:583-596.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    /* EdgeFunctionForward(p, a, b) = (p.x - a.x) * (b.y - a.y) - (p.y - a.y) * (b.x - a.x) */
    float t0 = (px - ax) * (by - ay);
    float t1 = (py - ay) * (bx - ax);
    return t0 - t1;
}

static inline float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
static inline float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

/* verts [N,V,3] f32 (already in pytorch3d NDC, i.e. after the wrapper's x,y negation renderer.py:403),
 * faces [N,F,3] i32, outputs pix_to_face [N,H,W] i64 (packed index n*F+f, -1 = empty),
 * zbuf [N,H,W] f32 (-1 empty), bary [N,H,W,3] f32 (-1 empty). */
void n3d_oracle_rasterize(const float* verts, const int32_t* faces, int N, int V, int F, int H, int W,
                          int64_t* pix_to_face, float* zbuf, float* bary) {
    const float kEps = 1e-8f;
    for (int64_t i = 0; i < (int64_t)N * H * W; ++i) { pix_to_face[i] = -1; zbuf[i] = -1.f; }
    for (int64_t i = 0; i < (int64_t)N * H * W * 3; ++i) bary[i] = -1.f;
    for (int n = 0; n < N; ++n) {
        const float* vb = verts + (int64_t)n * V * 3;
        for (int f = 0; f < F; ++f) {
            const int32_t* fi = faces + ((int64_t)n * F + f) * 3;
            const float x0 = vb[fi[0] * 3 + 0], y0 = vb[fi[0] * 3 + 1], z0 = vb[fi[0] * 3 + 2];
            const float x1 = vb[fi[1] * 3 + 0], y1 = vb[fi[1] * 3 + 1], z1 = vb[fi[1] * 3 + 2];
            const float x2 = vb[fi[2] * 3 + 0], y2 = vb[fi[2] * 3 + 1], z2 = vb[fi[2] * 3 + 2];
            const float xmin = fmin3(x0, x1, x2), xmax = fmax3(x0, x1, x2);
            const float ymin = fmin3(y0, y1, y2), ymax = fmax3(y0, y1, y2);
            const float zmax = fmax3(z0, z1, z2);
            if (zmax < 0.f) continue;
            const float face_area = edge_fn(x0, y0, x1, y1, x2, y2);     /* EdgeFunctionForward(v0, v1, v2) */
            if (face_area <= kEps && face_area >= -kEps) continue;       /* zero-area face */
            if (face_area < 0.f) continue;                                /* cull_backfaces */
            const float area = edge_fn(x2, y2, x0, y0, x1, y1) + kEps;    /* BarycentricCoordsForward denominator */
            /* conservative pixel range of the bbox (xf decreases with xi); exact test is done per pixel */
            int xi_lo = (int)floorf((1.f - xmax) * 0.5f * (float)W) - 2, xi_hi = (int)ceilf((1.f - xmin) * 0.5f * (float)W) + 2;
            int yi_lo = (int)floorf((1.f - ymax) * 0.5f * (float)H) - 2, yi_hi = (int)ceilf((1.f - ymin) * 0.5f * (float)H) + 2;
            if (xi_lo < 0) xi_lo = 0;
            if (yi_lo < 0) yi_lo = 0;
            if (xi_hi > W - 1) xi_hi = W - 1;
            if (yi_hi > H - 1) yi_hi = H - 1;
            for (int yi = yi_lo; yi <= yi_hi; ++yi) {
                const float yf = -1.f + (2.f * (float)(H - 1 - yi) + 1.f) / (float)H;
                if (yf < ymin || yf > ymax) continue;
                for (int xi = xi_lo; xi <= xi_hi; ++xi) {
                    const float xf = -1.f + (2.f * (float)(W - 1 - xi) + 1.f) / (float)W;
                    if (xf < xmin || xf > xmax) continue;
                    const float w0 = edge_fn(xf, yf, x1, y1, x2, y2) / area;
                    const float w1 = edge_fn(xf, yf, x2, y2, x0, y0) / area;
                    const float w2 = edge_fn(xf, yf, x0, y0, x1, y1) / area;
                    const float pz = (w0 * z0 + w1 * z1) + w2 * z2;
                    if (pz < 0.f) continue;
                    if (!(w0 > 0.f && w1 > 0.f && w2 > 0.f)) continue;   /* blur_radius 0: strictly inside only */
                    const int64_t p = ((int64_t)n * H + yi) * W + xi;
                    if (pix_to_face[p] < 0 || pz < zbuf[p]) {             /* nearest; ties keep the lower face index */
                        pix_to_face[p] = (int64_t)n * F + f;
                        zbuf[p] = pz;
                        bary[p * 3 + 0] = w0; bary[p * 3 + 1] = w1; bary[p * 3 + 2] = w2;
                    }
                }
            }
        }
    }
}

/* img [H,W] f32 in/out: pixels 4-connected to (0,0) whose value lies in [seed-lo, seed+up] are set to newval. */
void n3d_oracle_floodfill(float* img, int H, int W, float lo, float up, float newval) {
    const float seed = img[0];
    const float vmin = seed - lo, vmax = seed + up;
    uint8_t* mark = (uint8_t*)calloc((size_t)H * W, 1);
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)H * W);
    int sp = 0;
    stack[sp++] = 0; mark[0] = 1;
    while (sp > 0) {
        const int p = stack[--sp];
        const int y = p / W, x = p % W;
        const int nb[4][2] = {{y - 1, x}, {y + 1, x}, {y, x - 1}, {y, x + 1}};
        for (int k = 0; k < 4; ++k) {
            const int yy = nb[k][0], xx = nb[k][1];
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int q = yy * W + xx;
            if (mark[q]) continue;
            const float v = img[q];
            if (v >= vmin && v <= vmax) { mark[q] = 1; stack[sp++] = q; }
        }
    }
    for (int i = 0; i < H * W; ++i) if (mark[i]) img[i] = newval;
    free(mark); free(stack);
}
