// Forward pass of the B200 DIB-R rasterizer, fourth design ("v4"): no CTA barrier and no per-tile CTA anywhere.  The
// work is cut into two launches whose WARPS are independent of each other:
//
//   dibr_coverage_kernel  FACE-parallel hidden-surface pass into a global z-buffer (8 B per pixel, kept all-zero
//       between forwards).  A warp takes 32 consecutive faces; the front faces with a pixel centre in their bbox are
//       walked four at a time by 8 lanes each: a CHEAP conservative inside test (approximate reciprocal, a tolerance far
//       above its error) queues the (face, pixel) pairs that may be inside in a per-warp ring, and 32 queued pairs at a
//       time get the EXACT barycentric solve in the frozen fp32 order (two IEEE divisions) and one 64-bit atomicMax on
//       (orderable z << 32 | ~face id): the winner is the face with the largest z and, on ties, the smallest index --
//       what the reference's ascending loop with a strict '>' produces, independent of traversal order.  78 % of the
//       bbox tests miss and never reach the divisions.  Faces with more than 512 bbox pixels (listed by the set-up
//       kernel) are spread over a whole CTA each.
//   dibr_tiles_kernel     PIXEL-parallel: one warp per 8x4 pixel block of every touched 16x16 tile (work items fetched
//       heaviest tile first from a device counter), one warp per untouched tile (plain fill).  The warp reads (and
//       clears) its 32 z-buffer entries, interpolates the attributes of the covered pixels (128-bit loads), and, if it
//       has uncovered pixels, runs the soft silhouette on its own: it streams the tile's face bitmap (one bit per face,
//       set by the set-up kernel's exact binning) into ascending face ids, keeps the faces whose expanded pixel range
//       (in the record) meets an open pixel of the block in a ring, and 32 at a time transposes their 32-bit pixel
//       masks (5 shuffles) into per-pixel face masks in ascending face order, cuts them at the first K, turns them into a
//       flat (pixel, face) pair list the 32 lanes evaluate evenly, and lets each pixel fold its own results in face order.
//       Every output tensor is written exactly once, 128 bits per store (row segments of 8 pixels), through a
//       per-warp shared-memory transpose.
//
// Replaces kaolin v0.1's dr_cuda_forward_render_batch + dr_cuda_forward_prob_batch, which the reference calls at
// lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 (one thread per pixel looping over ALL faces,
// SURVEY.md 8(a) rows a6/a7).  Instead of the reference's five B x H x W x 30 scratch tensors (rasterizer.py:144-148)
// only the K-th accepted face id is kept, folded into imidx (see include/dibr_b200.h).  Earlier designs: v2 (one CTA
// per tile, kept as dibr_forward_v2.cu, DIBR_FWD_IMPL=2) and v3 (persistent tile CTAs with cp.async.bulk staging, in the
// history); profiles/r02_forward_v3_ab.md has the A/B that led here.
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

#ifdef DIBR_ITEM_TIMING
// debug: [0] longest item (cycles), [1] sum of item cycles, [2] items, [3] items above 20k cycles, [4] above 50k, [5] above 100k,
// [6] cycles of soft items, [7] soft items
__device__ unsigned long long g_dbg[8];
__device__ unsigned long long g_seg[16];
#define SEG_MARK(k) do { if ((threadIdx.x & 31) == 0) { const long long t_ = clock64(); atomicAdd(&g_seg[k], (unsigned long long)(t_ - t_seg)); t_seg = t_; } } while (0)
extern "C" void dibr_debug_item_cycles(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_dbg, sizeof(unsigned long long) * 8);
    cudaMemcpyFromSymbol(out8 + 8, g_seg, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_dbg, z, 64); cudaMemcpyToSymbol(g_seg, z, sizeof(z)); }
}
#else
#define SEG_MARK(k) do { } while (0)
#endif

constexpr int CW = 8;                           // warps per CTA, both kernels
constexpr int CQ = 64;                          // coverage: per-warp candidate ring (< 32 left + <= 32 new)
constexpr int BW = 8, BH = 4;                   // pixel block of one warp
constexpr int WGROUPS = 6;                      // soft: bitmap words are read 6 x 32 at a time (a LINEMOD-sized image in one batch)
constexpr int IDCAP = 128;                      // soft: face ids per flush (= ring capacity)
constexpr int QCAP = 256;                       // soft: pair list of a round
constexpr int BLOCKS_PER_TILE = (TILE / BW) * (TILE / BH);
static_assert(BLOCKS_PER_TILE == 8 && BW * BH == 32, "item = tile * 8 + block, lane = pixel");

// ================================================================================================================
// untouched tiles
// ================================================================================================================
// @phase fill untouched
// One WARP fills a full, 16 B-aligned 16x16 tile of one [H,W,CH] image with 128-bit stores: lanes walk whole rows
// (12 float4 per row for 3 channels, 4 for one), the row pointer advances by a constant.
__device__ __forceinline__ void fill_tile3_warp(float* __restrict__ img, int width, size_t pix0, float val)
{
    const int lane = threadIdx.x & 31;
    if (lane >= 24) return;
    const int r = (lane >= 12) ? 1 : 0, c = lane - 12 * r;
    const float4 v = make_float4(val, val, val, val);
    float4* p = reinterpret_cast<float4*>(img + (pix0 + (size_t)r * width) * 3) + c;
    const int step = (2 * width * 3) >> 2;               // two rows, in float4
#pragma unroll
    for (int k = 0; k < TILE / 2; k++) { *p = v; p += step; }
}
__device__ __forceinline__ void fill_tile1_warp(float* __restrict__ img, int width, size_t pix0, float val)
{
    const int lane = threadIdx.x & 31;
    const float4 v = make_float4(val, val, val, val);
    float4* p = reinterpret_cast<float4*>(img + pix0 + (size_t)(lane >> 2) * width) + (lane & 3);
    p[0] = v;
    p[(8 * width) >> 2] = v;
}

// one warp fills rows [0,th) x [0,tw) of one [H,W,ch] image tile with `val`
__device__ __forceinline__ void fill_tile_warp(float* __restrict__ img, int width, int ch, int tx0, int ty0, int tw, int th, float val, bool fast)
{
    if (fast && ch == 3) { fill_tile3_warp(img, width, (size_t)ty0 * width + tx0, val); return; }
    if (fast && ch == 1) { fill_tile1_warp(img, width, (size_t)ty0 * width + tx0, val); return; }
    const bool aligned = (((size_t)width * ch) & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (tw == TILE && th == TILE && aligned) {                     // tx0 is a multiple of 16: rows start 16 B aligned
        const int rv = 4 * ch;                                     // float4 per tile row
        const float4 v = make_float4(val, val, val, val);
        for (int i = threadIdx.x & 31; i < TILE * rv; i += 32) {
            const int r = i / rv;
            reinterpret_cast<float4*>(img + ((size_t)(ty0 + r) * width + tx0) * ch)[i - r * rv] = v;
        }
        return;
    }
    const int rowf = tw * ch;
    for (int i = threadIdx.x & 31; i < th * rowf; i += 32) {
        const int r = i / rowf;
        img[((size_t)(ty0 + r) * width + tx0) * ch + (i - r * rowf)] = val;
    }
}

// tile ids of the plan: image | tile row | tile column
__device__ __forceinline__ void unpack_tile(int packed, int& b, int& ty, int& tx) {
    b = (int)((unsigned)packed >> 20); ty = (packed >> 10) & 1023; tx = packed & 1023;
}

// nothing near this tile: zeros everywhere (imcomp = 1: empty product).  One warp per tile.
__device__ __forceinline__ void fill_untouched_warp(const FwdParams& P, int packed_tile)
{
    int b, ty, tx;
    unpack_tile(packed_tile, b, ty, tx);
    const int tx0 = tx * TILE, ty0 = ty * TILE;
    const int tw = min(TILE, P.width - tx0), th = min(TILE, P.height - ty0);
    const size_t img_pix = (size_t)b * P.height * P.width;
    // full tile, every tensor 16 B aligned with rows that keep the alignment
    const bool fast = tw == TILE && th == TILE && P.vec_out && (P.width & 3) == 0 &&
                      (((uintptr_t)P.improb | (uintptr_t)P.imidx | (uintptr_t)P.imcomp) & 15) == 0;
    if ((threadIdx.x & 31) == 0 && P.min_group >= 0 && f2ord(0.0f) < __ldcg(P.out_min)) atomicMin(P.out_min, f2ord(0.0f));
    for (int g = 0; g < P.n_out; g++) fill_tile_warp(P.out[g] + img_pix * P.out_ch[g], P.width, P.out_ch[g], tx0, ty0, tw, th, 0.0f, fast);
    fill_tile_warp(P.improb + img_pix, P.width, 1, tx0, ty0, tw, th, 0.0f, fast);
    fill_tile_warp(reinterpret_cast<float*>(P.imidx + img_pix), P.width, 1, tx0, ty0, tw, th, 0.0f, fast);
    fill_tile_warp(P.imcomp + img_pix, P.width, 1, tx0, ty0, tw, th, 1.0f, fast);
}

// ================================================================================================================
// coverage
// ================================================================================================================
// @phase cov exact
// exact coverage + depth test of one queued (face, pixel) pair
__device__ __forceinline__ void cov_exact(const FwdParams& P, unsigned g, unsigned xy)
{
    const float4* rp = reinterpret_cast<const float4*>(P.recs + g);
    const float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
    FaceRec r;
    r.ax = a.x; r.ay = a.y; r.bx = a.z; r.by = a.w; r.cx = b.x; r.cy = b.y;
    r.az = b.z; r.bz = b.w; r.cz = c.x;
    const FaceK fk = make_facek(r);
    const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
    float w0, w1, w2;
    if (!bary(fk, __ldg(P.xs + x), __ldg(P.ys + y), w0, w1, w2)) return;
    float z0 = blend(w0, w1, w2, fk.az, fk.bz, fk.cz);
    if (!(z0 > -1000.0f)) return;                 // "z0 <= znow" against the initial depth -1000
    z0 = z0 + 0.0f;                               // -0 -> +0 so equal depths compare equal
    const int bimg = __float_as_int(c.z);
    const int f_lo = P.face_offsets ? __ldg(P.face_offsets + bimg) : bimg * P.faces_per_image;
    const unsigned long long key = ((unsigned long long)f2ord(z0) << 32) | (unsigned long long)(0xffffffffu - (g - (unsigned)f_lo));
    atomicMax(P.zbuf + ((size_t)bimg * P.height + y) * P.width + x, key);
}

struct CovRing { unsigned int* qf; unsigned int* qp; int pn; };

// every lane calls: queue the lanes' candidate pairs, run the exact test when 32 are waiting
__device__ __forceinline__ void cov_push(const FwdParams& P, CovRing& R, bool cand, unsigned g, unsigned xy)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(full, cand);
    if (!bal) return;
    if (cand) {
        const int pos = R.pn + __popc(bal & ((1u << lane) - 1u));
        R.qf[pos] = g; R.qp[pos] = xy;
    }
    R.pn += __popc(bal);
    __syncwarp(full);
    if (R.pn >= 32) {
        cov_exact(P, R.qf[lane], R.qp[lane]);
        const int rest = R.pn - 32;
        unsigned mf = 0u, mp = 0u;
        if (lane < rest) { mf = R.qf[32 + lane]; mp = R.qp[32 + lane]; }
        __syncwarp(full);
        if (lane < rest) { R.qf[lane] = mf; R.qp[lane] = mp; }
        R.pn = rest;
        __syncwarp(full);
    }
}

// @phase cov cheap
// per-face constants of the conservative test
struct CheapFace { float ax, ay, m, p, n, q, rk3; bool sure; };
__device__ __forceinline__ CheapFace cheap_face(const float4 a, const float4 b) {
    CheapFace F;
    F.ax = a.x; F.ay = a.y;
    F.m = __fsub_rn(a.z, a.x); F.p = __fsub_rn(a.w, a.y); F.n = __fsub_rn(b.x, a.x); F.q = __fsub_rn(b.y, a.y);
    const float k3 = __fmaf_rn(F.m, F.q, -__fmul_rn(F.n, F.p));
    F.sure = fabsf(k3) >= 32.0f;                  // below that the reference's "+ 1e-15" matters: no shortcut
    F.rk3 = __frcp_rn(k3);
    return F;
}
// false only if the exact test is certain to reject the pixel
__device__ __forceinline__ bool cheap_inside(const CheapFace& F, float x0, float y0) {
    if (!F.sure) return true;
    const float sx = __fsub_rn(x0, F.ax), ty = __fsub_rn(y0, F.ay);
    const float k1 = __fmaf_rn(sx, F.q, -__fmul_rn(F.n, ty));
    const float k2 = __fmaf_rn(F.m, ty, -__fmul_rn(sx, F.p));
    const float w1 = k1 * F.rk3, w2 = k2 * F.rk3;                 // ~2e-7 relative from the exact quotients
    const float tol = 1e-4f * (1.0f + fabsf(w1) + fabsf(w2));
    return !(w1 < -tol || w2 < -tol || (1.0f - w1 - w2) < -tol);
}

// per-warp table of the 32 faces a warp walks in phase 1
struct CovFaces {
    int excl[32];                   // first flat pixel index of face l (faces without work have zero pixels)
    unsigned int c0nc[32];          // first column | columns << 16
    unsigned int r0[32];            // first row
    unsigned int magic[32];         // ceil(2^32 / columns)
    float4 fa[32], fb[32];          // ax ay m p | n q rk3 sure
};

__global__ void __launch_bounds__(CW * 32) dibr_coverage_kernel(const __grid_constant__ FwdParams P)
{
    __shared__ unsigned int s_qf[CW][CQ], s_qp[CW][CQ];
    __shared__ CovFaces s_faces[CW];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CovRing R;
    R.qf = s_qf[warp]; R.qp = s_qp[warp]; R.pn = 0;

    // ---- phase 1: 32 consecutive faces per warp.  Lane = face: the constants of the conservative test go to shared
    //      memory; then the bbox pixels of all 32 faces are walked as ONE flat index range, 32 pixels per turn, so
    //      every lane has a pixel whatever the faces' sizes are.
    const int gbase = (blockIdx.x * CW + warp) * 32;
    CovFaces& T = s_faces[warp];
    int npx = 0;
    if (gbase + lane < P.total_faces) {
        const FaceRec* rp = P.recs + gbase + lane;
        const float4* r4 = reinterpret_cast<const float4*>(rp);
        const float4 a = __ldg(r4), b = __ldg(r4 + 1);
        const float nz = __ldg(&rp->nz);
        const uint2 rg = __ldg(reinterpret_cast<const uint2*>(rp) + 6);           // cols, rows
        const int c0 = (int)(rg.x & 0xffffu), nc = (int)(rg.x >> 16) - c0;
        const int r0 = (int)(rg.y & 0xffffu), nr = (int)(rg.y >> 16) - r0;
        if (nz >= 0.0f && nc > 0 && nr > 0 && (long long)nc * nr <= BIG_FACE_PIXELS) {     // K1 culls normalz < 0
            npx = nc * nr;
            const CheapFace F = cheap_face(a, b);
            T.c0nc[lane] = (unsigned)c0 | ((unsigned)nc << 16);
            T.r0[lane] = (unsigned)r0;
            T.magic[lane] = (unsigned)((0x100000000ull + (unsigned)nc - 1ull) / (unsigned)nc);      // exact i / nc for i * nc < 2^32
            T.fa[lane] = make_float4(F.ax, F.ay, F.m, F.p);
            T.fb[lane] = make_float4(F.n, F.q, F.rk3, F.sure ? 1.0f : 0.0f);
        }
    }
    int incl = npx;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += t;
    }
    T.excl[lane] = incl - npx;
    const int total = __shfl_sync(full, incl, 31);
    __syncwarp(full);
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        bool cand = false;
        unsigned xy = 0u, fg = 0u;
        if (j < total) {
            int f = 0;                                  // last face whose range starts at or before j
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) if (T.excl[f + st] <= j) f += st;
            const int i = j - T.excl[f];
            const unsigned cn = T.c0nc[f];
            const int nc = (int)(cn >> 16);
            const int row = (nc > 1) ? (int)__umulhi((unsigned)i, T.magic[f]) : i;
            const int x = (int)(cn & 0xffffu) + (i - row * nc), y = (int)T.r0[f] + row;
            const float4 fa = T.fa[f], fb = T.fb[f];
            CheapFace F;
            F.ax = fa.x; F.ay = fa.y; F.m = fa.z; F.p = fa.w; F.n = fb.x; F.q = fb.y; F.rk3 = fb.z; F.sure = fb.w != 0.0f;
            cand = cheap_inside(F, __ldg(P.xs + x), __ldg(P.ys + y));
            xy = (unsigned)x | ((unsigned)y << 16);
            fg = (unsigned)(gbase + f);
        }
        cov_push(P, R, cand, fg, xy);
    }
    // ---- phase 2: the large faces (listed by the set-up kernel), one CTA each: rows over the warps, columns over the lanes
    const int nbig = min(__ldg(P.big_count), P.total_faces);
    for (int k = blockIdx.x; k < nbig; k += gridDim.x) {
        const unsigned fg = (unsigned)__ldg(P.big_list + k);
        const float4* rp = reinterpret_cast<const float4*>(P.recs + fg);
        const CheapFace F = cheap_face(__ldg(rp), __ldg(rp + 1));
        const uint2 rg = __ldg(reinterpret_cast<const uint2*>(P.recs + fg) + 6);
        const int c0 = (int)(rg.x & 0xffffu), c1 = (int)(rg.x >> 16), r0 = (int)(rg.y & 0xffffu), r1 = (int)(rg.y >> 16);
        for (int y = r0 + warp; y < r1; y += CW) {
            const float y0 = __ldg(P.ys + y);
            for (int xb = c0; xb < c1; xb += 32) {
                const int x = xb + lane;
                const bool cand = (x < c1) && cheap_inside(F, __ldg(P.xs + x), y0);
                cov_push(P, R, cand, fg, (unsigned)x | ((unsigned)y << 16));
            }
        }
    }
    if (lane < R.pn) cov_exact(P, R.qf[lane], R.qp[lane]);
    // ---- phase 3: the untouched tiles (plan bucket 0) are plain fills that depend on nothing: whoever is done with its faces
    //      takes them from a device counter, so their HBM traffic overlaps the latency-bound phases above
    // @phase fill untouched
    const int n_empty = __ldg(P.order_cnt);
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(P.list_counts + 3, 1);
        j = __shfl_sync(full, j, 0);
        if (j >= n_empty) break;
        fill_untouched_warp(P, __ldg(P.order_seg + j));
    }
}

// ================================================================================================================
// tiles: resolve + soft silhouette, one warp per 8x4 block
// ================================================================================================================
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// @phase soft round
// rows = lanes, columns = bits: returns column `lane` of the 32x32 bit matrix as this lane's word
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
    unsigned m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((y & ~m) >> j) | (x & ~m)) : ((x & m) | ((y & m) << j));
        m ^= (m << (j >> 1));        // 0000ffff -> 00ff00ff -> 0f0f0f0f -> 33333333 -> 55555555
    }
    return x;
}

struct SoftState {
    float q, cc;        // running 1 - prod(1-p) and prod(1-p)
    int c;              // faces accepted so far
    int kth;            // image-local id of the K-th accepted face, or -1
    bool open;          // uncovered pixel inside the image
};

struct BlockCtx {
    int x0, y0;                     // first pixel of the block
    int f_lo;                       // first face of the image
    float zscale, sentinel;
    int knum;
    unsigned int* u;                // [QCAP] face ids from the bitmap waiting for the range test (IDCAP of them); during the rounds: the pair list
    unsigned int* ring_m;           // [IDCAP] block-pixel masks of the faces that meet open pixels of the block, ascending ids
    unsigned int* ring_f;           // [IDCAP] their image-local ids
    float4* cor_a;                  // [2][32] corners ax ay bx by of the round's 32 faces (double buffered: the next round's are in flight)
    float2* cor_b;                  // [2][32] cx cy
    unsigned int* wbuf;             // [WGROUPS * 32] bitmap words of the tile (masked for this block), WGROUPS groups at a time
    float* sxy;                     // [8 + 4] pixel-centre x of the block's columns, y of its rows
    unsigned char* soft_flag;       // [total_faces] faces that entered some soft product (the backward's work list is built from it)
#ifdef DIBR_ITEM_TIMING
    unsigned dbg_info;              // flushes << 24 | ids << 12 | ring entries of the last flush
#endif
    unsigned open32;                // pixels that still accept faces
};

__device__ __forceinline__ float soft_pair(const BlockCtx& B, int buf, int j, int p) {
    const float4 a = B.cor_a[buf * 32 + j];
    const float2 d = B.cor_b[buf * 32 + j];
    const SoftHit h = soft_distance(a.x, a.y, a.z, a.w, d.x, d.y, B.sxy[p & 7], B.sxy[8 + (p >> 3)], B.sentinel);
    return soft_prob_enc(h.d2 * B.zscale);
}
__device__ __forceinline__ void soft_fold(float v, SoftState& st) {
    float p, om;
    soft_prob_dec(v, p, om);
    st.q = fmaf(p, st.cc, st.q);                    // 1 - prod(1-p), accurate for small p
    st.cc = st.cc * om;                             // prod(1-p), accurate for p near 1
}

// One round: ring entries [r0, r0 + n), n <= 32, in ascending face order; their corners are in corner buffer `buf`.
// Everything it reads is in shared memory.
__device__ __forceinline__ void soft_round(BlockCtx& B, SoftState& st, int r0, int n, int buf)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned int* __restrict__ Q = B.u;
    const unsigned m32 = (lane < n) ? (B.ring_m[r0 + lane] & B.open32) : 0u;
    unsigned tm = transpose32(m32, lane);           // lane = pixel: bit j <-> ring entry r0 + j holds this pixel
    int nb = __popc(tm);
    if (st.c + nb > B.knum) {                       // first-K rule: keep the lowest K - c entries
        const int keep = B.knum - st.c;
        tm = (keep > 0) ? (tm & ((2u << __fns(tm, 0u, keep)) - 1u)) : 0u;
        nb = max(keep, 0);
    }
    if (nb > 0 && st.c + nb == B.knum) st.kth = (int)B.ring_f[r0 + 31 - __clz(tm)];     // the K-th accepted face closes the pixel
    st.c += nb;
    // the faces that entered some pixel's product: flagged for the backward (plain byte stores, everybody writes 1)
    if ((__reduce_or_sync(full, tm) >> lane) & 1u) B.soft_flag[B.f_lo + (int)B.ring_f[r0 + lane]] = 1;
    int incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(full, incl, 31);
    const int excl = incl - nb;
    if (total > 0 && total <= QCAP) {
        // flat pair list, pixel-major, faces ascending inside a pixel
        int pos = excl;
        for (unsigned t = tm; t; t &= t - 1) Q[pos++] = ((unsigned)lane << 8) | (unsigned)(__ffs(t) - 1);
        __syncwarp(full);
        for (int e = lane; e < total; e += 32) {    // evaluated evenly
            const unsigned ent = Q[e];
            Q[e] = __float_as_uint(soft_pair(B, buf, (int)(ent & 0xffu), (int)(ent >> 8)));
        }
        __syncwarp(full);
        for (int k = 0; k < nb; k++) soft_fold(__uint_as_float(Q[excl + k]), st);
        __syncwarp(full);
    } else if (total > QCAP) {
        // dense round (large faces over the whole block): every pixel has many entries, lane = pixel is balanced
        for (unsigned t = tm; t; t &= t - 1) soft_fold(soft_pair(B, buf, __ffs(t) - 1, lane), st);
        __syncwarp(full);
    }
    B.open32 = __ballot_sync(full, st.open && st.c < B.knum);
}

// block-pixel mask of an expanded pixel range
__device__ __forceinline__ unsigned range_mask(const BlockCtx& B, uint2 er) {
    const int lo = max((int)(er.x & 0xffffu), B.x0) - B.x0, hi = min((int)(er.x >> 16), B.x0 + BW) - B.x0;
    const int rlo = max((int)(er.y & 0xffffu), B.y0) - B.y0, rhi = min((int)(er.y >> 16), B.y0 + BH) - B.y0;
    if (hi <= lo || rhi <= rlo) return 0u;
    const unsigned cm = (1u << hi) - (1u << lo), rm = (1u << rhi) - (1u << rlo);
    return (cm * 0x01010101u) & (((rm * 0x00204081u) & 0x01010101u) * 0xffu);
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// request the corners of ring entries [r0, r0 + 32) into corner buffer `buf` (asynchronous copies straight into shared memory)
__device__ __forceinline__ void request_corners(const FwdParams& P, const BlockCtx& B, int r0, int n, int buf) {
    const int lane = threadIdx.x & 31;
    if (r0 + lane < n) {
        const float4* rp = reinterpret_cast<const float4*>(P.recs + B.f_lo + B.ring_f[r0 + lane]);
        cp_async16(B.cor_a + buf * 32 + lane, rp);
        cp_async8(B.cor_b + buf * 32 + lane, rp + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// @phase soft flush
// The `nid` face ids waiting in B.u: range test for all of them (one batch of gathers per 128 ids), then the rounds of the
// faces that meet the block, 32 at a time, with the next round's corners in flight.
__device__ __forceinline__ void soft_flush(const FwdParams& P, BlockCtx& B, SoftState& st, int nid)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
#ifdef DIBR_ITEM_TIMING
    long long t_seg = clock64();
#endif
    int n = 0;
    static_assert(IDCAP == 128, "one batch of four gathers per lane");
    {
        const int cb = 0;
        uint2 er[4];
        unsigned fid[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = cb + 32 * c + lane;
            fid[c] = (i < nid) ? B.u[i] : 0u;
            er[c] = (i < nid) ? __ldg(P.fbox + B.f_lo + fid[c]) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const unsigned m = (cb + 32 * c + lane < nid) ? (range_mask(B, er[c]) & B.open32) : 0u;
            const unsigned bal = __ballot_sync(full, m != 0u);
            if (m) {
                const int pos = n + __popc(bal & lt);
                B.ring_m[pos] = m; B.ring_f[pos] = fid[c];
            }
            n += __popc(bal);
        }
    }
    __syncwarp(full);
#ifdef DIBR_ITEM_TIMING
    B.dbg_info = ((B.dbg_info & 0xff000000u) + (1u << 24)) | ((unsigned)nid << 12) | (unsigned)n;
#endif
    SEG_MARK(8);
    if (n == 0) return;
    request_corners(P, B, 0, n, 0);
    int buf = 0;
    for (int r0 = 0; r0 < n && B.open32; r0 += 32, buf ^= 1) {
        if (r0 + 32 < n) {
            request_corners(P, B, r0 + 32, n, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp(full);
        SEG_MARK(9);
        soft_round(B, st, r0, min(32, n - r0), buf);
        SEG_MARK(10);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // a round may have closed the last pixel with a request in flight
    __syncwarp(full);
}

// @phase soft walk
// Soft silhouette of one block: the tile's face bitmap is read WGROUPS x 32 words at a time (one batch of loads), masked
// with the per-word block masks and expanded into ascending face ids; every IDCAP ids (or at the end) a flush.  One loop
// with a single call site per stage (the code is large: copies would thrash the instruction cache).
__device__ __forceinline__ void soft_block(const FwdParams& P, BlockCtx& B, SoftState& st, const uint32_t* __restrict__ words,
                                           const unsigned char* __restrict__ wmask, int blk, int nw, int id0, uint32_t word0)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ngroups = (nw + 31) >> 5;
#ifdef DIBR_ITEM_TIMING
    long long t_seg = clock64();
#endif
    int nid = 0;                                    // ids waiting in B.u
    int g = -1;                                     // group held in `word` (one word per lane)
    uint32_t word = 0u;
    while (B.open32) {
        const unsigned left = __ballot_sync(full, word != 0u);              // lanes that still hold bits of this group
        const bool input_done = (left == 0u) && (g + 1 >= ngroups);
        if (nid == IDCAP || (input_done && nid > 0)) {
            SEG_MARK(11);
            soft_flush(P, B, st, nid);
            nid = 0;
#ifdef DIBR_ITEM_TIMING
            t_seg = clock64();
#endif
        } else if (input_done) {
            break;
        } else if (left == 0u) {
            g++;
            if (g % WGROUPS == 0) {                 // next WGROUPS groups: all loads first, then the masks
                uint32_t wd[WGROUPS];
                unsigned mk[WGROUPS];
#pragma unroll
                for (int k = 0; k < WGROUPS; k++) {
                    const int w = 32 * (g + k) + lane;
                    wd[k] = (w < nw) ? __ldg(words + w) : 0u;
                    mk[k] = (w < nw) ? (unsigned)__ldg(wmask + w) : 0u;
                }
#pragma unroll
                for (int k = 0; k < WGROUPS; k++) B.wbuf[32 * k + lane] = ((mk[k] >> blk) & 1u) ? wd[k] : 0u;     // words without a face near this block are skipped
                if (g == 0) B.wbuf[lane] = word0;   // requested (and masked) at the top of the item
            }
            word = B.wbuf[32 * (g % WGROUPS) + lane];                   // each lane reads back what it wrote
        } else {
            const int cnt = __popc(word);
            int pos, total;
            if (__ballot_sync(full, cnt > 1) == 0u) {       // usual case (faces of a tile are spread over the words): at most one bit per lane
                pos = __popc(left & ((1u << lane) - 1u));
                total = __popc(left);
            } else {
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(full, incl, o);
                    if (lane >= o) incl += t;
                }
                total = __shfl_sync(full, incl, 31);
                pos = incl - cnt;
            }
            const int room = IDCAP - nid;           // > 0: a full list is flushed first
            const int id = id0 + ((32 * g + lane) << 5);
            while (word && pos < room) {            // lane-major = ascending ids; what does not fit waits for the next turn
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                B.u[nid + pos] = (unsigned)(id + bit);
                pos++;
            }
            nid += min(total, room);
            __syncwarp(full);
        }
    }
}

// 65536 / d + 1: (i * c_inv[d]) >> 16 == i / d for the small i used here
__constant__ unsigned c_inv32[33] = {0u, 65537u, 32769u, 21846u, 16385u, 13108u, 10923u, 9363u, 8193u, 7282u, 6554u, 5958u, 5462u, 5042u, 4682u, 4370u, 4097u,
                                     3856u, 3641u, 3450u, 3277u, 3121u, 2979u, 2850u, 2731u, 2622u, 2521u, 2428u, 2341u, 2260u, 2185u, 2115u, 2049u};

// @phase block resolve
// One touched-tile block: z-buffer -> attributes -> soft silhouette -> outputs.  Warp-uniform control flow.
// `rowsel` >= 0: the item is ONE pixel row of the block (the blocks of the heaviest tiles are split four ways so that no
// single warp carries a whole dense block: the longest item sets the kernel's duration).
__device__ __forceinline__ void block_item(const FwdParams& P, int4 desc, int blk, int rowsel, float* __restrict__ stage, BlockCtx& B)
{
#ifdef DIBR_ITEM_TIMING
    long long t_seg = clock64();
#endif
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    int b, tile_y, tile_x;
    unpack_tile(desc.x, b, tile_y, tile_x);
    const int W = P.width, H = P.height, D = P.num_attr;
    B.x0 = tile_x * TILE + (blk & 1) * BW;
    B.y0 = tile_y * TILE + (blk >> 1) * BH;
    const int x = B.x0 + (lane & 7), y = B.y0 + (lane >> 3);
    const bool valid = (x < W) && (y < H) && (rowsel < 0 || (lane >> 3) == rowsel);
    const int f_lo = desc.y, nw = desc.z;
    B.f_lo = f_lo;
    const size_t img_pix = (size_t)b * H * W;
    const size_t px = img_pix + (size_t)y * W + x;
    const uint32_t* words = P.bins + desc.w;                    // the tile's face bitmap and its per-word block masks
    const unsigned char* wmask = P.wordmask + desc.w;
    // ---- everything that depends on the descriptor alone is requested here, before any of it is used (the SM issues in
    //      order: a load placed behind the first use of another one waits a full round trip longer)
    unsigned long long key = 0ull;
    if (valid) key = __ldcg(P.zbuf + px);
    const unsigned tblocks = __ldg(P.tile_blocks + (size_t)b * tiles_x * tiles_y + tile_y * tiles_x + tile_x);
    uint32_t word0 = 0u;
    unsigned wm0 = 0u;
    if (P.knum > 0 && lane < nw) { word0 = __ldg(words + lane); wm0 = (unsigned)__ldg(wmask + lane); }
    unsigned omin = 0u;
    if (P.min_group >= 0) omin = __ldcg(P.out_min);
    // ---- this block's z-buffer entries; they go back to zero for the next forward
    if (key != 0ull) P.zbuf[px] = 0ull;
    const int fw = (key != 0ull) ? (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) : -1;
    const unsigned covered32 = __ballot_sync(full, fw >= 0);
    if (!((tblocks >> blk) & 1u) || !((wm0 >> blk) & 1u)) word0 = 0u;
    SEG_MARK(0);
    // ---- resolve: winner's weights, attribute interpolation.  Requests first: the winner's record and attribute rows,
    //      and (soft phase) the rest of the tile's bitmap
    float v[DIBR_MAX_ATTR_INTERNAL];
#pragma unroll
    for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) v[d] = 0.f;
    SoftState st;
    st.q = 0.f; st.cc = 1.f; st.c = 0; st.kth = -1;
    st.open = valid && fw < 0;
    B.open32 = (P.knum > 0 && ((tblocks >> blk) & 1u)) ? __ballot_sync(full, st.open) : 0u;      // no face's expanded range meets the block: no soft phase
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 c1 = make_float2(0.f, 0.f);
    const float* a = P.face_attr + (size_t)(f_lo + max(fw, 0)) * 3 * D;
    if (fw >= 0) {
        const float4* rp = reinterpret_cast<const float4*>(P.recs + f_lo + fw);
        c0 = __ldg(rp);
        c1 = __ldg(reinterpret_cast<const float2*>(rp + 1));
        prefetch_l1(a);
        prefetch_l1(a + 3 * D - 1);
    }
    if (B.open32) {
#pragma unroll
        for (int k = 1; k < WGROUPS; k++)
            if (32 * k + lane < nw) { prefetch_l1(words + 32 * k + lane); if ((lane & 3) == 0) prefetch_l1(wmask + 32 * k + lane); }
    }
    if (fw >= 0) {
        FaceRec r;
        r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
        r.az = r.bz = r.cz = 0.f;
        const FaceK fk = make_facek(r);
        float w0, w1, w2;
        bary(fk, __ldg(P.xs + x), __ldg(P.ys + y), w0, w1, w2);
        if ((D & 3) == 0) {                  // the three corner rows as 128-bit loads
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d += 4) {
                if (d < D) {
                    const float4 r0 = __ldg(reinterpret_cast<const float4*>(a + d));
                    const float4 r1 = __ldg(reinterpret_cast<const float4*>(a + D + d));
                    const float4 r2 = __ldg(reinterpret_cast<const float4*>(a + 2 * D + d));
                    v[d] = blend(w0, w1, w2, r0.x, r1.x, r2.x);
                    v[d + 1] = blend(w0, w1, w2, r0.y, r1.y, r2.y);
                    v[d + 2] = blend(w0, w1, w2, r0.z, r1.z, r2.z);
                    v[d + 3] = blend(w0, w1, w2, r0.w, r1.w, r2.w);
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
                if (d < D) v[d] = blend(w0, w1, w2, __ldg(a + d), __ldg(a + D + d), __ldg(a + 2 * D + d));
        }
    }
    // ---- batch-global minimum of one output group (normal map), zeros of the uncovered pixels included
    if (P.min_group >= 0) {
        float vmin = 3.0e38f;
#pragma unroll
        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
            if (d < D && valid && ((P.min_mask >> d) & 1u)) vmin = fminf(vmin, v[d]);
        const unsigned ov = __reduce_min_sync(full, f2ord(vmin));
        if (lane == 0 && ov != f2ord(3.0e38f) && ov < omin) atomicMin(P.out_min, ov);
    }
    SEG_MARK(1);
    // ---- the attribute images: row segments of 8 pixels, 128 bits per store, through a per-warp transpose
    // @phase block write
    if (P.vec_out && rowsel < 0 && B.x0 + BW <= W && B.y0 + BH <= H) {
        if (covered32) {
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
                if (d < D) stage[P.chan_off32[d] + lane * P.chan_stride[d]] = v[d];
            __syncwarp(full);
        }
        int off = 0;
        const size_t pix0 = img_pix + (size_t)B.y0 * W + B.x0;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < P.n_out; g++) {
            const int ch = P.out_ch[g];
            float* __restrict__ base = P.out[g] + pix0 * ch;
            if (ch == 3) {                          // 24 float4: lane -> row lane / 6, column lane % 6
                if (lane < 24) {
                    const int row = (lane * 43) >> 8;                       // lane / 6 for lane < 24
                    const float4 val = covered32 ? reinterpret_cast<const float4*>(stage + off)[lane] : zero4;
                    reinterpret_cast<float4*>(base + (size_t)row * W * 3)[lane - row * 6] = val;
                }
            } else if (ch == 1) {                   // 8 float4: lane -> row lane / 2, column lane % 2
                if (lane < 8) {
                    const float4 val = covered32 ? reinterpret_cast<const float4*>(stage + off)[lane] : zero4;
                    reinterpret_cast<float4*>(base + (size_t)(lane >> 1) * W)[lane & 1] = val;
                }
            } else {
                const int rv = 2 * ch;                                      // float4 per row segment
                const unsigned inv = c_inv32[min(rv, 32)];
                for (int i = lane; i < 4 * rv; i += 32) {
                    const int row = (rv <= 32) ? (int)(((unsigned)i * inv) >> 16) : i / rv;
                    const float4 val = covered32 ? reinterpret_cast<const float4*>(stage + off)[i] : zero4;
                    reinterpret_cast<float4*>(base + (size_t)row * W * ch)[i - row * rv] = val;
                }
            }
            off += 32 * ch;
        }
        __syncwarp(full);
    } else if (valid) {
#pragma unroll
        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
            if (d < D) P.chan_out[d][px * P.chan_stride[d]] = v[d];
    }
    SEG_MARK(2);
    // ---- winners are flagged for the backward's colour work list (plain byte stores; run-length de-duplicated along the row segment)
    // @phase block lists
    if (covered32) {
        const int prev = __shfl_up_sync(full, fw, 1);
        if (fw >= 0 && ((lane & 7) == 0 || prev != fw)) reinterpret_cast<unsigned char*>(P.face_flags)[f_lo + fw] = 1;
    }
    {   // uncovered pixels of the block, one byte per row, for the backward's soft part
        const unsigned ub = __ballot_sync(full, st.open);
        if (lane < BH && (rowsel < 0 || lane == rowsel) && B.y0 + lane < H && B.x0 < W)
            P.open8[((size_t)b * H + B.y0 + lane) * ((W + 7) >> 3) + (B.x0 >> 3)] = (unsigned char)((ub >> (8 * lane)) & 0xffu);
    }
    // ---- soft silhouette of the uncovered pixels
    if (B.open32) {
        if (lane < BW) B.sxy[lane] = (B.x0 + lane < W) ? __ldg(P.xs + B.x0 + lane) : 0.f;
        else if (lane < BW + BH) B.sxy[lane] = (B.y0 + lane - BW < H) ? __ldg(P.ys + B.y0 + lane - BW) : 0.f;
        __syncwarp(full);
        soft_block(P, B, st, words, wmask, blk, nw, ((f_lo >> 5) << 5) - f_lo, word0);
    }
    SEG_MARK(3);
    // @phase block write
    if (valid) {
        const size_t gp = px;
        if (fw >= 0) { P.improb[gp] = 1.0f; P.imcomp[gp] = 0.0f; P.imidx[gp] = fw + 1; }
        else {
            P.improb[gp] = fminf(st.q, 1.0f);           // the recurrence can overshoot 1 by an ulp
            P.imcomp[gp] = st.cc;
            P.imidx[gp] = (st.kth >= 0) ? -(st.kth + 1) : 0;
        }
    }
    SEG_MARK(4);
}

#ifndef DIBR_TILES_CTAS_PER_SM
#define DIBR_TILES_CTAS_PER_SM 4
#endif
#ifndef DIBR_ITEM_CHUNK
#define DIBR_ITEM_CHUNK 1
#endif
constexpr int ITEM_CHUNK = DIBR_ITEM_CHUNK;      // work items per counter atomic
#ifndef DIBR_HEAVY_BUCKET
#define DIBR_HEAVY_BUCKET 31
#endif
constexpr int HEAVY_BUCKET = DIBR_HEAVY_BUCKET;  // tiles that list more than 32 * (HEAVY_BUCKET - 1) faces are cut into pixel rows
static_assert(BH == 4 && HEAVY_BUCKET >= 1 && HEAVY_BUCKET <= 31, "item decoding");
constexpr int WARP_WORDS = 384 + 2 * IDCAP + QCAP + WGROUPS * 32 + 16;       // per-warp scratch (4.4 KB: shared memory a CTA takes is L1 cache its loads lose)
// @phase prologue
__global__ void __launch_bounds__(CW * 32, DIBR_TILES_CTAS_PER_SM)
dibr_tiles_kernel(const __grid_constant__ FwdParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_bstart[34];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    const int ntiles = tiles_x * tiles_y * P.batch;
    // per-warp scratch: corner buffers, ring (masks, ids), id / pair list, bitmap words, pixel centres, the transpose buffer
    const int warp_words = WARP_WORDS;
    unsigned int* const wbase = reinterpret_cast<unsigned int*>(smem_raw) + (size_t)warp * warp_words;
    BlockCtx B;
    B.cor_a = reinterpret_cast<float4*>(wbase);                                     // 16 B aligned: first
    B.cor_b = reinterpret_cast<float2*>(wbase + 256);
    B.ring_m = wbase + 384; B.ring_f = B.ring_m + IDCAP;
    B.u = B.ring_f + IDCAP;
    B.wbuf = B.u + QCAP;
    B.sxy = reinterpret_cast<float*>(B.wbuf + WGROUPS * 32);
    float* const stage = reinterpret_cast<float*>(B.ring_m);                        // the transpose buffer is dead before the soft phase starts
    static_assert(2 * IDCAP + QCAP >= 32 * DIBR_MAX_ATTR_INTERNAL, "transpose buffer aliases ring + pair list");
    B.soft_flag = reinterpret_cast<unsigned char*>(P.face_flags) + P.total_faces;
    B.zscale = (float)P.delta / ((float)P.multiplier * (float)P.multiplier);
    B.sentinel = 4.0f * (float)P.multiplier * (float)P.multiplier;
    B.knum = P.knum;
    B.open32 = 0u; B.x0 = B.y0 = B.f_lo = 0;
    // ---- the plan (set-up: plan_tiles_kernel): tiles by cost bucket.  bstart[l] = first position of bucket 31-l in
    //      heaviest-first order; bucket 0 (empty bitmaps) comes last.
    if (warp == 0) {
        static_assert(ORDER_BUCKETS == 32, "one bucket per lane");
        const int n = __ldg(P.order_cnt + (ORDER_BUCKETS - 1 - lane));
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        s_bstart[lane] = incl - n;
        if (lane == 31) s_bstart[32] = incl;
    }
    __syncthreads();                                    // the only CTA barrier of the forward pass
    const int n_touched = s_bstart[31];
    const int n_heavy = s_bstart[ORDER_BUCKETS - HEAVY_BUCKET];               // tiles of the buckets >= HEAVY_BUCKET come first
    const int n_heavy_items = n_heavy * BLOCKS_PER_TILE * BH;                  // ... and are cut into single pixel rows
    const int n_items = n_heavy_items + (n_touched - n_heavy) * BLOCKS_PER_TILE;
    const int my_start = s_bstart[lane];
    // ---- work items from a device counter, ITEM_CHUNK at a time: the blocks of the touched tiles, heaviest tile first.
    //      The next chunk is requested while the current one is processed.
    // position in heaviest-first order -> descriptor
    auto desc_of = [&](int ti) {
        const unsigned le = __ballot_sync(full, my_start <= ti);       // lane 31 (bucket 0) starts at n_touched > ti
        const int l = 31 - __clz(le);
        return __ldg(P.tile_desc + (size_t)(ORDER_BUCKETS - 1 - l) * ntiles + (ti - __shfl_sync(full, my_start, l)));
    };
    auto decode = [&](int item, int& ti, int& blk, int& rowsel) {
        if (item < n_heavy_items) { ti = item >> 5; blk = (item >> 2) & 7; rowsel = item & 3; }
        else { const int j = item - n_heavy_items; ti = n_heavy + (j >> 3); blk = j & 7; rowsel = -1; }
    };
    // software pipeline over the items: while item i is processed, the descriptor of item i+1 and the index of item i+2 are in flight
    int item = 0, item1 = 0;
    if (lane == 0) { item = atomicAdd(P.list_counts + 2, 1); item1 = atomicAdd(P.list_counts + 2, 1); }
    item = __shfl_sync(full, item, 0);
    item1 = __shfl_sync(full, item1, 0);
    int ti, blk, rowsel;
    int4 desc = make_int4(0, 0, 0, 0);
    if (item < n_items) { decode(item, ti, blk, rowsel); desc = desc_of(ti); }
    while (item < n_items) {
        int item2 = 0;
        if (lane == 0) item2 = atomicAdd(P.list_counts + 2, 1);
        int ti1 = 0, blk1 = 0, rowsel1 = -1;
        int4 desc1 = make_int4(0, 0, 0, 0);
        if (item1 < n_items) { decode(item1, ti1, blk1, rowsel1); desc1 = desc_of(ti1); }
#ifdef DIBR_ITEM_TIMING
        const long long t_item = clock64();
        B.dbg_info = 0u;
#endif
        block_item(P, desc, blk, rowsel, stage, B);
#ifdef DIBR_ITEM_TIMING
        if (lane == 0) {
            const unsigned long long dt = (unsigned long long)(clock64() - t_item);
            atomicMax(&g_dbg[0], (dt << 32) | B.dbg_info); atomicAdd(&g_dbg[1], dt); atomicAdd(&g_dbg[2], 1ull);
            if ((B.dbg_info >> 24) >= 2u) { atomicAdd(&g_seg[5], dt); atomicAdd(&g_seg[6], 1ull); }
            if (dt > 20000ull) atomicAdd(&g_dbg[3], 1ull);
            if (dt > 50000ull) atomicAdd(&g_dbg[4], 1ull);
            if (dt > 100000ull) atomicAdd(&g_dbg[5], 1ull);
            if (dt > 6000ull) { atomicAdd(&g_dbg[6], dt); atomicAdd(&g_dbg[7], 1ull); }
        }
#endif
        item = item1; desc = desc1; blk = blk1; rowsel = rowsel1;
        item1 = __shfl_sync(full, item2, 0);
    }
}

// ================================================================================================================
// out = (n - min) / (||n - min|| + 1e-5) * mask  (renderer_dibr.py:284-285).  Four pixels per thread: three 128-bit
// loads of normals, one of the mask, three 128-bit stores (HBM-bound: 28 B per pixel).
__device__ __forceinline__ void normal_map_pixel(float a, float b, float c, float m, float mn, float& oa, float& ob, float& oc) {
    a -= mn; b -= mn; c -= mn;
    const float len = sqrtf(a * a + b * b + c * c) + 1e-5f;
    oa = a / len * m; ob = b / len * m; oc = c / len * m;
}
__global__ void __launch_bounds__(256) normal_map_kernel(const float* __restrict__ n, const float* __restrict__ mask,
                                                         const unsigned int* __restrict__ min_ordered, float* __restrict__ out, long long npix, int vec_ok)
{
    const float mn = ord2f(*min_ordered);
    const long long nquad = vec_ok ? (npix >> 2) : 0;
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long q = t0; q < nquad; q += stride) {
        const float4* np4 = reinterpret_cast<const float4*>(n) + 3 * q;
        const float4 x = __ldcs(np4), y = __ldcs(np4 + 1), z = __ldcs(np4 + 2);
        const float4 m = __ldcs(reinterpret_cast<const float4*>(mask) + q);
        float4 ox, oy, oz;
        normal_map_pixel(x.x, x.y, x.z, m.x, mn, ox.x, ox.y, ox.z);
        normal_map_pixel(x.w, y.x, y.y, m.y, mn, ox.w, oy.x, oy.y);
        normal_map_pixel(y.z, y.w, z.x, m.z, mn, oy.z, oy.w, oz.x);
        normal_map_pixel(z.y, z.z, z.w, m.w, mn, oz.y, oz.z, oz.w);
        float4* op4 = reinterpret_cast<float4*>(out) + 3 * q;
        op4[0] = ox; op4[1] = oy; op4[2] = oz;
    }
    for (long long i = (nquad << 2) + t0; i < npix; i += stride) {
        float oa, ob, oc;
        normal_map_pixel(n[3 * i], n[3 * i + 1], n[3 * i + 2], mask[i], mn, oa, ob, oc);
        out[3 * i] = oa; out[3 * i + 1] = ob; out[3 * i + 2] = oc;
    }
}

int launch_normal_map(const float* n, const float* mask, const unsigned int* min_ordered, float* out, long long npix, cudaStream_t stream)
{
    if (npix == 0) return 0;
    const int vec_ok = (((uintptr_t)n | (uintptr_t)mask | (uintptr_t)out) & 15) == 0;
    const long long work = vec_ok ? (npix + 3) / 4 : npix;
    const int grid = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
    normal_map_kernel<<<grid, 256, 0, stream>>>(n, mask, min_ordered, out, npix, vec_ok);
    return (int)cudaGetLastError();
}

// forward implementation switch for the A/B (profiles/): the tile-CTA kernel of dibr_forward_v2.cu is the default (it is the
// faster one, profiles/r02_forward_ab.md); DIBR_FWD_IMPL=4 in the environment selects the barrier-free design of this file
int forward_impl() {
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("DIBR_FWD_IMPL");
        impl = (e && e[0] == '4') ? 4 : 2;
    }
    return impl;
}

int launch_forward(const FwdParams& P, cudaStream_t stream)
{
    if (forward_impl() == 2) return launch_forward_v2(P, stream);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    // per-device launch configuration, looked up once
    static int sms_of[64];
    if (sms_of[dev] == 0) {
        int sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(dibr_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(unsigned) * CW * WARP_WORDS));
        if (e != cudaSuccess) return (int)e;
        sms_of[dev] = sms;
    }
    {
        // one warp per 32 faces; at least a few CTAs per SM so the fills of the untouched tiles have the whole machine
        const long long ntl = (long long)((P.width + TILE - 1) / TILE) * ((P.height + TILE - 1) / TILE) * P.batch;
        int cgrid = (P.total_faces + CW * 32 - 1) / (CW * 32);
        const int fgrid = (int)(ntl / CW + 1 < (long long)sms_of[dev] * 4 ? ntl / CW + 1 : (long long)sms_of[dev] * 4);
        if (cgrid < fgrid) cgrid = fgrid;
        dibr_coverage_kernel<<<cgrid, CW * 32, 0, stream>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    const size_t smem = sizeof(unsigned) * CW * (size_t)WARP_WORDS;
    const long long ntiles = (long long)((P.width + TILE - 1) / TILE) * ((P.height + TILE - 1) / TILE) * P.batch;
    const long long want = (ntiles * BLOCKS_PER_TILE + CW - 1) / CW;
    const int grid = (int)(want < (long long)sms_of[dev] * DIBR_TILES_CTAS_PER_SM ? want : (long long)sms_of[dev] * DIBR_TILES_CTAS_PER_SM);
    dibr_tiles_kernel<<<grid, CW * 32, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
