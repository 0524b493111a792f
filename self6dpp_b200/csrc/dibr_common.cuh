// Shared device helpers for the B200 DIB-R kernels.  Every arithmetic step that decides
// coverage (and therefore the bit-exact face-index buffer) is written with explicit
// round-to-nearest intrinsics so nvcc can neither contract nor reorder it; the order is the one
// frozen in oracle/dibr_oracle_body.h (which restates kaolin v0.1's dr_cuda_forward_render_batch,
// reached from /root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dibr {

constexpr int TILE = 16;            // forward CTA tile (pixels per side): one pixel per thread
#ifndef DIBR_FWD_THREADS
#define DIBR_FWD_THREADS 256
#endif
constexpr int FWD_THREADS = DIBR_FWD_THREADS;   // threads per tile CTA (8 warps, one 8x4 pixel block each in the soft phase)
#ifndef DIBR_LCAP
#define DIBR_LCAP 512
#endif
constexpr int LCAP = DIBR_LCAP;     // faces per in-shared-memory batch of a tile (multiple of 16, <= 512)
constexpr int BIGCAP = 32;          // deferred large faces per batch
#ifndef DIBR_BIG_AREA
#define DIBR_BIG_AREA 64
#endif
constexpr int BIG_AREA = DIBR_BIG_AREA;   // pixels of a face inside the tile above which the CTA cooperates
static_assert(FWD_THREADS == TILE * TILE, "one pixel per thread");
constexpr int MAX_IMAGE_SIDE = 16384;           // largest image side the ABI accepts
constexpr int BIG_FACE_PIXELS = 512;            // bbox pixel centres above which a front face is rasterised by the whole grid

// 64 B face record, written by the set-up kernels.
struct __align__(16) FaceRec {
    float ax, ay, bx, by;           // 2D corners, already x multiplier
    float cx, cy, az, bz;           // third corner, view-space z of a and b
    float cz, nz, image, local_id;  // view-space z of c, z of the face normal (< 0: back face), image index and the face's id inside its image (int bits)
    // pixel ranges, lo | hi << 16 (hi exclusive, both clamped to the image): the pixel centres inside the bbox of the 2D
    // corners (rasterizer.py:49-52, half-open) and inside the expanded bbox (rasterizer.py:54-57).  All four are 0
    // (empty) for a face with a non-finite corner.
    unsigned int cols, rows, ecols, erows;
};

// float -> unsigned with the same ordering
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// (float)((double)num / ((double)den + 1e-15)), bit-exact.  For |den| >= 32 the double sum
// den + 1e-15 rounds back to den (half an ulp of a double in [32,64) is 3.6e-15) and a double
// quotient rounded to float equals the correctly rounded float quotient (53 >= 2*24+2), so the
// fast path is the IEEE fp32 division.
__device__ __forceinline__ float div_eps(float num, float den) {
    if (fabsf(den) >= 32.0f) return __fdiv_rn(num, den);
    return (float)((double)num / ((double)den + 1e-15));
}

// per-face constants of the barycentric solve
struct FaceK {
    float ax, ay, m, p, n, q, k3;
    float az, bz, cz;
};

__device__ __forceinline__ FaceK make_facek(const FaceRec& r) {
    FaceK k;
    k.ax = r.ax; k.ay = r.ay;
    k.m = __fsub_rn(r.bx, r.ax); k.p = __fsub_rn(r.by, r.ay);
    k.n = __fsub_rn(r.cx, r.ax); k.q = __fsub_rn(r.cy, r.ay);
    k.k3 = __fmaf_rn(k.m, k.q, -__fmul_rn(k.n, k.p));
    k.az = r.az; k.bz = r.bz; k.cz = r.cz;
    return k;
}

// barycentric weights of pixel centre (x0,y0); returns false when outside (any weight < 0)
__device__ __forceinline__ bool bary(const FaceK& k, float x0, float y0, float& w0, float& w1, float& w2) {
    const float s = __fsub_rn(x0, k.ax), t = __fsub_rn(y0, k.ay);
    const float k1 = __fmaf_rn(s, k.q, -__fmul_rn(k.n, t));
    const float k2 = __fmaf_rn(k.m, t, -__fmul_rn(s, k.p));
    w1 = div_eps(k1, k.k3);
    w2 = div_eps(k2, k.k3);
    w0 = __fsub_rn(__fsub_rn(1.0f, w1), w2);
    return !(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f);
}

__device__ __forceinline__ float blend(float w0, float w1, float w2, float r0, float r1, float r2) {
    return __fmaf_rn(w2, r2, __fmaf_rn(w0, r0, __fmul_rn(w1, r1)));
}

// pixel-centre coordinates: "1.0 * multiplier / width * (2*w + 1 - width)" in double, rounded once
__device__ __forceinline__ float pix_x(int w, int width, int multiplier) {
    return (float)(1.0 * multiplier / width * (2 * w + 1 - width));
}
__device__ __forceinline__ float pix_y(int h, int height, int multiplier) {
    return (float)(1.0 * multiplier / height * (height - 2 * h - 1));
}

// first column c in [0,W] whose pixel centre pix_x(c) >= x.  Centre c sits at f == c; the fp32 guess is off by ~1e-5
// columns (2e-6 |f| for the largest images), so only an x within `tol` of a centre is decided by the exact value.
__device__ __forceinline__ int first_col_ge(float x, int W, int M) {
    const float f = fmaf(x, 0.5f * (float)W / (float)M, 0.5f * (float)(W - 1));
    if (!(f > -1.0f)) return 0;                   // left of everything (or NaN)
    if (!(f < (float)W)) return W;
    const float cf = ceilf(f);
    int c = (int)cf;
    const float d = cf - f, tol = 2e-3f + 2e-6f * fabsf(f);
    if (d < tol || d > 1.0f - tol) {
        const int k = min(max((int)rintf(f), 0), W - 1);
        c = (pix_x(k, W, M) >= x) ? k : k + 1;
    }
    return min(max(c, 0), W);
}
// first row r in [0,H] whose pixel centre pix_y(r) < y  (centres descend: centre r sits at g == r)
__device__ __forceinline__ int first_row_lt(float y, int H, int M) {
    const float g = fmaf(-y, 0.5f * (float)H / (float)M, 0.5f * (float)(H - 1));
    if (!(g > -1.0f)) return 0;                   // above everything (or NaN)
    if (!(g < (float)H)) return H;
    const float ff = floorf(g);
    int r = (int)ff + 1;
    const float d = g - ff, tol = 2e-3f + 2e-6f * fabsf(g);
    if (d < tol || d > 1.0f - tol) {
        const int k = min(max((int)rintf(g), 0), H - 1);
        r = (pix_y(k, H, M) < y) ? k : k + 1;
    }
    return min(max(r, 0), H);
}

// Soft-silhouette distance of a pixel to a face: min over the 3 edges (perpendicular distance when
// the foot lies on the segment, else the sentinel 4 m^2) and the 3 corners, first minimum wins
// (edge0, edge1, edge2, corner0, corner1, corner2).  Written in local differences, which is better
// conditioned than the reference's A x + B y + C form; compared with the fp64 oracle at 1e-5.
struct SoftHit {
    float d2;       // squared distance, multiplier units
    int kase;       // 0..5
    float cr, len2; // edge cases: cross product and squared edge length
};

__device__ __forceinline__ SoftHit soft_distance(float x1, float y1, float x2, float y2, float x3, float y3,
                                                 float x0, float y0, float sentinel) {
    const float dx1 = x0 - x1, dy1 = y0 - y1;
    const float dx2 = x0 - x2, dy2 = y0 - y2;
    const float dx3 = x0 - x3, dy3 = y0 - y3;
    SoftHit h;
    h.d2 = sentinel; h.kase = 0; h.cr = 0.f; h.len2 = 1.f;
    float best = 3.0e38f;
    // edges i -> i+1
#define DIBR_EDGE(I, DX, DY, XA, YA, XB, YB)                                   \
    {                                                                            \
        const float ex = (XB) - (XA), ey = (YB) - (YA);                          \
        const float len2 = fmaf(ex, ex, ey * ey);                                \
        const float dot = fmaf((DX), ex, (DY) * ey);                             \
        const float cr = fmaf((DX), ey, -((DY) * ex));                           \
        const bool on = (dot >= 0.0f) && (dot <= len2) && (len2 > 0.0f);         \
        const float d = on ? __fdividef(cr * cr, len2) : sentinel;               \
        if (best > d) { best = d; h.kase = (I); h.cr = cr; h.len2 = len2; }      \
    }
    DIBR_EDGE(0, dx1, dy1, x1, y1, x2, y2)
    DIBR_EDGE(1, dx2, dy2, x2, y2, x3, y3)
    DIBR_EDGE(2, dx3, dy3, x3, y3, x1, y1)
#undef DIBR_EDGE
    const float v1 = fmaf(dx1, dx1, dy1 * dy1);
    const float v2 = fmaf(dx2, dx2, dy2 * dy2);
    const float v3 = fmaf(dx3, dx3, dy3 * dy3);
    if (best > v1) { best = v1; h.kase = 3; }
    if (best > v2) { best = v2; h.kase = 4; }
    if (best > v3) { best = v3; h.kase = 5; }
    h.d2 = best;
    return h;
}

// prob = exp(-z) and om = 1 - prob, both with full relative accuracy, carried by ONE float: a value with the sign
// bit clear is om (small z: alternating series, truncation < 2e-9 relative), one with the sign bit set is -prob.
__device__ __forceinline__ float soft_prob_enc(float z) {
    if (z < 0.25f) {
        float t = fmaf(z, -1.0f / 7.0f, 1.0f);
        t = fmaf(z * (-1.0f / 6.0f), t, 1.0f);
        t = fmaf(z * (-1.0f / 5.0f), t, 1.0f);
        t = fmaf(z * (-1.0f / 4.0f), t, 1.0f);
        t = fmaf(z * (-1.0f / 3.0f), t, 1.0f);
        t = fmaf(z * (-1.0f / 2.0f), t, 1.0f);
        return z * t;
    }
    return __uint_as_float(__float_as_uint(expf(-z)) | 0x80000000u);
}
__device__ __forceinline__ void soft_prob_dec(float v, float& p, float& om) {
    if (__float_as_int(v) < 0) { p = -v; om = 1.0f - p; }
    else { om = v; p = 1.0f - om; }
}
__device__ __forceinline__ void soft_prob(float z, float& p, float& om) {
    soft_prob_dec(soft_prob_enc(z), p, om);
}

}  // namespace dibr
