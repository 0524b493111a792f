// Forward pass of the B200 DIB-R rasterizer, third design ("v3"): PERSISTENT CTAs (256 threads = the 256 pixels of a
// 16x16 screen tile) that walk the plan the set-up call left -- touched tiles heaviest first, then the untouched tiles
// (one warp fills one) -- so the grid is (SMs x resident CTAs), not (tiles).  Per touched tile:
//
//   A  list     read the tile's face bitmap (one bit per face of the image, set by the set-up kernel's binning), scan the
//               popcounts, expand the set bits into the ascending list of face ids; STAGE the listed faces' records
//               (48 B each: corners, corner depths, normal z) into shared memory with one bulk async copy per face
//               (cp.async.bulk + mbarrier expect-tx; -DDIBR_GATHER=0/1 select plain loads / cp.async for the A/B);
//               per face: clip its bbox and its expanded bbox to the tile (column / row masks), front faces with a
//               pixel centre in range go on the raster list.
//   B  coverage two passes.  CHEAP: 8 lanes per raster face walk its pixels with a conservative sign test (approximate
//               reciprocal, a tolerance far above its error) and queue the (face, pixel) pairs that may be inside;
//               EXACT: one thread per queued pair does the barycentric solve in the frozen fp32 order and a 64-bit
//               shared-memory atomicMax on (orderable z | ~rank).  78 % of the bbox tests miss: they never reach the
//               two IEEE divisions.  The winner is the face with the largest z and, on ties, the smallest index --
//               what the reference's ascending loop with a strict '>' produces, independent of traversal order.
//   C  resolve  one pixel per thread: winner's weights, attribute interpolation (128-bit loads), everything goes to a
//               shared-memory copy of the tile's outputs.
//   D  soft     silhouette probability of the uncovered pixels, one 8x4 pixel block per WARP, no CTA barrier inside:
//               the faces whose expanded bbox meets an open pixel of the block are compacted (ballot) into a ring, 32
//               at a time their 32-bit pixel masks are transposed (5 shuffles) into per-pixel face masks in ascending
//               face order, cut at the first K, turned into a flat (pixel, face) pair list that the 32 lanes evaluate
//               evenly, and each pixel folds its own results in face order.
//   E  write    the tile's outputs leave shared memory as 128-bit coalesced stores, every tensor written once.
//
// Replaces kaolin v0.1's dr_cuda_forward_render_batch + dr_cuda_forward_prob_batch, which the reference calls at
// lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 (one thread per pixel looping over ALL faces,
// SURVEY.md 8(a) rows a6/a7).  Instead of the reference's five B x H x W x 30 scratch tensors (rasterizer.py:144-148)
// only the K-th accepted face id is kept, folded into imidx (see include/dibr_b200.h).  The previous design (one CTA
// per tile, 20 barriers per tile, exact test for every bbox pixel) is kept as dibr_forward_v2.cu for the A/B.
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

#ifdef DIBR_PHASE_TIMING
__device__ unsigned long long g_phase[8];
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(t_ - t_phase)); t_phase = t_; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

#ifndef DIBR_GATHER
#define DIBR_GATHER 2               // 0: ld.global + st.shared, 1: cp.async (LDGSTS), 2: cp.async.bulk + mbarrier (UBLKCP)
#endif
#ifndef DIBR_FWD_CTAS_PER_SM
#define DIBR_FWD_CTAS_PER_SM 4
#endif

constexpr int NT = FWD_THREADS;                 // 256 threads: one per pixel of the tile
constexpr int NWARP = NT / 32;
constexpr int BW = 8, BH = 4;                   // pixel block of one warp in the soft phase
constexpr int NBX = TILE / BW;
constexpr int CQ = 1024;                        // coverage candidates per round
constexpr int QCAP = 256;                       // soft (pixel, face) pairs per warp per round
constexpr int RING = 64;                        // compacted faces waiting for a soft round (< 32 left + <= 32 new)
constexpr int REC_V4 = 3;                       // float4 per staged record (the first 48 B of FaceRec)
static_assert(NBX * (TILE / BH) == NWARP, "one 8x4 block per warp");
static_assert(LCAP >= 32 && LCAP <= 512 && LCAP % 16 == 0 && TILE == 16, "rlist packing: 9-bit list index, 4-bit pixel coordinates");

struct __align__(16) FwdSmem {
    float4 rec[LCAP * REC_V4];                  // staged records: ax ay bx by | cx cy az bz | cz nz image -
    int lid[LCAP];                              // local face ids, ascending
    unsigned int smask[LCAP];                   // tile columns (bits 0-15) / rows (16-31) inside the face's expanded bbox
    union {
        struct {
            unsigned long long zkey[TILE * TILE];   // z-buffer
            unsigned int wf[TILE * TILE];       // winner's face id, committed per batch when a tile needs several
            unsigned int rlist[LCAP];           // raster candidates: list index | c0 | nc-1 | r0 | nr-1
            unsigned int candq[CQ];             // (list index << 8) | pixel
        } cov;
        unsigned int softq[NWARP][QCAP];        // per warp: (pixel << 16 | list index), then the encoded probability
    } u;
    unsigned int ring_m[NWARP][RING];           // per warp: block-pixel masks of the compacted faces
    unsigned short ring_li[NWARP][RING];        //           and their list indices
    unsigned char cnt[TILE * TILE];             // 255 = covered
    unsigned char soft_used[LCAP];              // listed faces that entered some pixel's soft product
    float xs[TILE], ys[TILE];
    int warp_tot[2][NWARP];
    int bstart[34];                             // plan: first position of bucket 31-l in heaviest-first order
    int lcount, rcount, qcount, pad0;
    unsigned long long mbar;                    // completion of the record copies (DIBR_GATHER == 2)
};

// @phase A gather
// ---- small PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// @phase A prep
// first column c in [0,n] with xs[c] >= x (xs ascending, pitch 1/inv_dx).  The arithmetic guess is exact unless x sits
// within 2e-3 columns of a pixel centre (its error is ~1e-5 columns); only then the table decides.
__device__ __forceinline__ int col_first_ge(const float* xs, int n, float x, float inv_dx) {
    const float f = (x - xs[0]) * inv_dx;
    const float cf = ceilf(f);
    int c = (int)fminf(fmaxf(cf, 0.f), (float)n);
    const float d = cf - f;
    if ((d < 2e-3f || d > 0.998f) && f > -2.0f && f < (float)(TILE + 2)) {
        while (c > 0 && xs[c - 1] >= x) c--;
        while (c < n && xs[c] < x) c++;
    }
    return c;
}
// first row r in [0,n] with ys[r] < y (ys descending)
__device__ __forceinline__ int row_first_lt(const float* ys, int n, float y, float inv_dy) {
    const float f = (ys[0] - y) * inv_dy;
    const float ff = floorf(f);
    int r = (int)fminf(fmaxf(ff + 1.0f, 0.f), (float)n);
    const float d = f - ff;
    if ((d < 2e-3f || d > 0.998f) && f > -2.0f && f < (float)(TILE + 2)) {
        while (r > 0 && ys[r - 1] < y) r--;
        while (r < n && ys[r] >= y) r++;
    }
    return r;
}

// @phase A expand
struct TileGeom {
    int tw, th;
    float inv_dx, inv_dy;
    const uint32_t* words;          // this tile's face bitmap (set-up kernel: bin_face)
    int nw;                         // its length in 32-face words
    int id0;                        // local face id of bit 0 of word 0  (<= 0)
};

// Phase A, first half.  Expands the tile's face bitmap from word `wpos` on into the ascending list of local face ids,
// until the list is full (a batch always ends on a word boundary).  Returns the first unread word; leaves s.lcount and
// zeroes s.rcount / s.qcount.  Uniform across the CTA; ends with a barrier.
__device__ int expand_batch(FwdSmem& s, int wpos, const TileGeom& T, int& parity)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lcount = 0;
    while (wpos < T.nw) {
        const int w = wpos + tid;
        uint32_t word = (w < T.nw) ? __ldg(T.words + w) : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s.warp_tot[parity][warp] = incl;
        __syncthreads();
        int base = lcount, total = 0;
#pragma unroll
        for (int k = 0; k < NWARP; k++) {
            const int t = s.warp_tot[parity][k];
            if (k < warp) base += t;
            total += t;
        }
        parity ^= 1;      // double buffered: a writer of the same parity two rounds on has passed the next round's barrier
        const int gi = base + incl;                             // list length after this thread's word
        bool fits = true;
        int taken = NT;                                         // words of this round that go on the list
        if (lcount + total > LCAP) {                            // the prefix of words that still fits (monotone in tid)
            fits = (gi <= LCAP);
            taken = __syncthreads_count(fits ? 1 : 0);
        }
        if (fits) {
            int slot = gi - cnt;
            const int id = T.id0 + (w << 5);
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                s.lid[slot++] = id + bit;
            }
        }
        if (taken < NT) {
            if (taken == 0) break;
            if (tid == taken - 1) s.lcount = gi;                // the thread of the last taken word knows the new length
            __syncthreads();
            lcount = s.lcount;
            wpos += taken;
            break;
        }
        lcount += total;
        wpos += NT;
    }
    wpos = min(wpos, T.nw);
    __syncthreads();
    if (tid == 0) { s.lcount = lcount; s.rcount = 0; s.qcount = 0; }
    __syncthreads();
    return wpos;
}

// @phase A gather
// Phase A, second half: stage the listed faces' records into shared memory.  Ends with the records visible to the CTA.
__device__ __forceinline__ void gather_records(FwdSmem& s, const FaceRec* __restrict__ recs, unsigned& mbar_parity)
{
    const int tid = threadIdx.x;
    const int lcount = s.lcount;
#if DIBR_GATHER == 2
    // one bulk async copy per face, completion counted in bytes by the mbarrier
    if (lcount > 0) {
        if (tid == 0) mbar_expect_tx(&s.mbar, (unsigned)lcount * (unsigned)(REC_V4 * 16));
        __syncthreads();                                        // expect-tx before any complete-tx
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of rec[] before the async writes
        for (int i = tid; i < lcount; i += NT) bulk_g2s(&s.rec[i * REC_V4], recs + s.lid[i], REC_V4 * 16, &s.mbar);
        mbar_wait(&s.mbar, mbar_parity);
        mbar_parity ^= 1u;
    }
#elif DIBR_GATHER == 1
    for (int i = tid; i < lcount; i += NT) {
        const float4* rp = reinterpret_cast<const float4*>(recs + s.lid[i]);
#pragma unroll
        for (int k = 0; k < REC_V4; k++) cp_async16(&s.rec[i * REC_V4 + k], rp + k);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
#else
    for (int i = tid; i < lcount; i += NT) {
        const float4* rp = reinterpret_cast<const float4*>(recs + s.lid[i]);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
        s.rec[i * REC_V4] = r0; s.rec[i * REC_V4 + 1] = r1; s.rec[i * REC_V4 + 2] = r2;
    }
    __syncthreads();
#endif
}

// @phase A prep
// Phase A, third part: per listed face the tile columns / rows inside its expanded bbox (smask) and, for front faces
// with a pixel centre in range, a raster-list entry.  Ends with a barrier.
__device__ void prep_faces(FwdSmem& s, const TileGeom& T, float ex, bool raster)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int lcount = s.lcount;
    for (int i0 = 0; i0 < lcount; i0 += NT) {
        const int i = i0 + tid;
        bool keep = false;
        unsigned int packed = 0u;
        if (i < lcount) {
            const float4 a = s.rec[i * REC_V4], b = s.rec[i * REC_V4 + 1];
            const float xmin = fminf(a.x, fminf(a.z, b.x)), xmax = fmaxf(a.x, fmaxf(a.z, b.x));     // rasterizer.py:49-52
            const float ymin = fminf(a.y, fminf(a.w, b.y)), ymax = fmaxf(a.y, fmaxf(a.w, b.y));
            {
                const int e0 = col_first_ge(s.xs, T.tw, xmin - ex, T.inv_dx), e1 = col_first_ge(s.xs, T.tw, xmax + ex, T.inv_dx);   // rasterizer.py:54-57
                const int q0 = row_first_lt(s.ys, T.th, ymax + ex, T.inv_dy), q1 = row_first_lt(s.ys, T.th, ymin - ex, T.inv_dy);
                s.smask[i] = (e1 > e0 && q1 > q0) ? (((1u << e1) - (1u << e0)) | (((1u << q1) - (1u << q0)) << 16)) : 0u;
            }
            s.soft_used[i] = 0;
            if (raster && s.rec[i * REC_V4 + 2].y >= 0.0f) {    // front face (K1 culls normalz < 0)
                const int c0 = col_first_ge(s.xs, T.tw, xmin, T.inv_dx), c1 = col_first_ge(s.xs, T.tw, xmax, T.inv_dx);
                const int q0 = row_first_lt(s.ys, T.th, ymax, T.inv_dy), q1 = row_first_lt(s.ys, T.th, ymin, T.inv_dy);
                if (c1 > c0 && q1 > q0) {
                    keep = true;
                    packed = (unsigned)i | ((unsigned)c0 << 9) | ((unsigned)(c1 - c0 - 1) << 13) |
                             ((unsigned)q0 << 17) | ((unsigned)(q1 - q0 - 1) << 21);
                }
            }
        }
        if (raster) {
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (bal) {
                const int leader = __ffs(bal) - 1;
                int rb = 0;
                if (lane == leader) rb = atomicAdd(&s.rcount, __popc(bal));
                rb = __shfl_sync(0xffffffffu, rb, leader);
                if (keep) s.u.cov.rlist[rb + __popc(bal & lt)] = packed;
            }
        }
    }
    __syncthreads();
}

// @phase B exact
__constant__ unsigned c_inv16[17] = {0u, 65537u, 32769u, 21846u, 16385u, 13108u, 10923u, 9363u, 8193u, 7282u, 6554u, 5958u, 5462u, 5042u, 4682u, 4370u, 4097u};

struct RasterEntry { int li, c0, nc, r0, nr; };
__device__ __forceinline__ RasterEntry unpack_entry(unsigned int p) {
    RasterEntry e;
    e.li = p & 511; e.c0 = (p >> 9) & 15; e.nc = ((p >> 13) & 15) + 1; e.r0 = (p >> 17) & 15; e.nr = ((p >> 21) & 15) + 1;
    return e;
}

__device__ __forceinline__ FaceK facek_from_list(const FwdSmem& s, int li) {
    const float4 a = s.rec[li * REC_V4], b = s.rec[li * REC_V4 + 1];
    FaceRec r;
    r.ax = a.x; r.ay = a.y; r.bx = a.z; r.by = a.w; r.cx = b.x; r.cy = b.y;
    r.az = b.z; r.bz = b.w; r.cz = s.rec[li * REC_V4 + 2].x;
    return make_facek(r);
}

// exact coverage test + depth test of one (face, pixel) pair
__device__ __forceinline__ void raster_pixel(FwdSmem& s, const FaceK& fk, unsigned rank, int lx, int ly) {
    float w0, w1, w2;
    if (!bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2)) return;
    float z0 = blend(w0, w1, w2, fk.az, fk.bz, fk.cz);
    if (!(z0 > -1000.0f)) return;                 // "z0 <= znow" against the initial depth -1000
    z0 = z0 + 0.0f;                               // -0 -> +0 so equal depths compare equal
    const unsigned long long key = ((unsigned long long)f2ord(z0) << 32) | (unsigned long long)(0xffffffffu - rank);
    atomicMax(&s.u.cov.zkey[ly * TILE + lx], key);
}

// @phase B cheap
// Phase B.  `nprev`: faces listed by earlier batches of this tile (ranks keep ascending across batches).
__device__ void coverage(FwdSmem& s, int nprev)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int rcount = s.rcount;
    const int g = tid >> 3, ql = tid & 7;                       // 32 groups of 8 lanes, 4 per warp
    // ---- cheap pass: which (face, pixel) pairs can be inside at all
    for (int e0 = 0; e0 < rcount; e0 += NT / 8) {
        const int e = e0 + g;
        RasterEntry en;
        en.li = 0; en.c0 = 0; en.nc = 1; en.r0 = 0; en.nr = 0;
        if (e < rcount) en = unpack_entry(s.u.cov.rlist[e]);
        const int npx = en.nc * en.nr;
        const int maxit = __reduce_max_sync(full, (npx + 7) >> 3);
        const float4 a = s.rec[en.li * REC_V4], b = s.rec[en.li * REC_V4 + 1];
        const float m = __fsub_rn(a.z, a.x), p = __fsub_rn(a.w, a.y), n = __fsub_rn(b.x, a.x), q = __fsub_rn(b.y, a.y);
        const float k3 = __fmaf_rn(m, q, -__fmul_rn(n, p));
        const bool sure = fabsf(k3) >= 32.0f;                   // below that the reference's "+ 1e-15" matters: no shortcut
        const float rk3 = __frcp_rn(k3);
        const unsigned inv = c_inv16[en.nc];                    // 65536 / nc + 1: exact i / nc for i < 256
        for (int it = 0; it < maxit; it++) {
            const int i = it * 8 + ql;
            bool cand = false;
            unsigned entry = 0u;
            if (i < npx) {
                const int row = (int)(((unsigned)i * inv) >> 16);
                const int lx = en.c0 + (i - row * en.nc), ly = en.r0 + row;
                cand = true;
                if (sure) {
                    const float sx = __fsub_rn(s.xs[lx], a.x), ty = __fsub_rn(s.ys[ly], a.y);
                    const float k1 = __fmaf_rn(sx, q, -__fmul_rn(n, ty));
                    const float k2 = __fmaf_rn(m, ty, -__fmul_rn(sx, p));
                    const float w1 = k1 * rk3, w2 = k2 * rk3;                   // ~2e-7 relative from the exact quotients
                    const float tol = 1e-4f * (1.0f + fabsf(w1) + fabsf(w2));
                    cand = !(w1 < -tol || w2 < -tol || (1.0f - w1 - w2) < -tol);
                }
                entry = ((unsigned)en.li << 8) | (unsigned)(ly * TILE + lx);
            }
            const unsigned bal = __ballot_sync(full, cand);
            if (bal) {
                const int leader = __ffs(bal) - 1;
                int qb = 0;
                if (lane == leader) qb = atomicAdd(&s.qcount, __popc(bal));
                qb = __shfl_sync(full, qb, leader) + __popc(bal & lt);
                if (cand && qb < CQ) s.u.cov.candq[qb] = entry;
            }
        }
    }
    __syncthreads();
    // @phase B exact
    // ---- exact pass over the queue, every lane busy
    const int nq = s.qcount;
    for (int j = tid; j < min(nq, CQ); j += NT) {
        const unsigned entry = s.u.cov.candq[j];
        const int li = (int)(entry >> 8), pix = (int)(entry & 255u);
        const FaceK fk = facek_from_list(s, li);
        raster_pixel(s, fk, (unsigned)(nprev + li), pix & (TILE - 1), pix >> 4);
    }
    if (nq > CQ) {
        // the queue overflowed (many layers of large faces): exact test for every bbox pixel of every raster face; what
        // the queue already delivered is delivered again, which an atomicMax does not mind
        for (int e = g; e < rcount; e += NT / 8) {
            const RasterEntry en = unpack_entry(s.u.cov.rlist[e]);
            const int npx = en.nc * en.nr;
            const FaceK fk = facek_from_list(s, en.li);
            const unsigned inv = c_inv16[en.nc];
            for (int i = ql; i < npx; i += 8) {
                const int row = (int)(((unsigned)i * inv) >> 16);
                raster_pixel(s, fk, (unsigned)(nprev + en.li), en.c0 + (i - row * en.nc), en.r0 + row);
            }
        }
    }
    __syncthreads();
}

// @phase D round
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

struct SoftGeom { int bx, by; float zscale, sentinel; int knum; };

__device__ __forceinline__ float soft_pair(const FwdSmem& s, const SoftGeom& G, int lj, int p) {
    const float4 a = s.rec[lj * REC_V4];
    const float4 d = s.rec[lj * REC_V4 + 1];
    const SoftHit h = soft_distance(a.x, a.y, a.z, a.w, d.x, d.y, s.xs[G.bx + (p & 7)], s.ys[G.by + (p >> 3)], G.sentinel);
    return soft_prob_enc(h.d2 * G.zscale);
}
__device__ __forceinline__ void soft_fold(float v, SoftState& st) {
    float p, om;
    soft_prob_dec(v, p, om);
    st.q = fmaf(p, st.cc, st.q);                    // 1 - prod(1-p), accurate for small p
    st.cc = st.cc * om;                             // prod(1-p), accurate for p near 1
}

// One round of the soft phase of one warp: the first `n` (<= 32) ring entries, in ascending face order.
__device__ __forceinline__ void soft_round(FwdSmem& s, const SoftGeom& G, SoftState& st, int n, unsigned& open32)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned short* ring_li = s.ring_li[warp];
    unsigned int* Q = s.u.softq[warp];
    const unsigned m32 = (lane < n) ? (s.ring_m[warp][lane] & open32) : 0u;
    unsigned tm = transpose32(m32, lane);           // lane = pixel: bit j <-> ring entry j holds this pixel
    int nb = __popc(tm);
    if (st.c + nb > G.knum) {                       // first-K rule: keep the lowest K - c entries
        const int keep = G.knum - st.c;
        tm = (keep > 0) ? (tm & ((2u << __fns(tm, 0u, keep)) - 1u)) : 0u;
        nb = max(keep, 0);
    }
    if (nb > 0 && st.c + nb == G.knum) st.kth = s.lid[ring_li[31 - __clz(tm)]];     // the K-th accepted face closes the pixel
    st.c += nb;
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
        for (unsigned t = tm; t; t &= t - 1) Q[pos++] = ((unsigned)lane << 16) | (unsigned)ring_li[__ffs(t) - 1];
        __syncwarp(full);
        for (int e = lane; e < total; e += 32) {    // evaluated evenly
            const unsigned ent = Q[e];
            const int lj = (int)(ent & 0xffffu);
            s.soft_used[lj] = 1;                    // benign race: everybody writes 1
            Q[e] = __float_as_uint(soft_pair(s, G, lj, (int)(ent >> 16)));
        }
        __syncwarp(full);
        for (int k = 0; k < nb; k++) soft_fold(__uint_as_float(Q[excl + k]), st);
        __syncwarp(full);
    } else if (total > QCAP) {
        // dense round (large faces over the whole block): every pixel has many entries, lane = pixel is balanced
        for (unsigned t = tm; t; t &= t - 1) {
            const int lj = (int)ring_li[__ffs(t) - 1];
            s.soft_used[lj] = 1;
            soft_fold(soft_pair(s, G, lj, lane), st);
        }
        __syncwarp(full);
    }
    open32 = __ballot_sync(full, st.open && st.c < G.knum);
}

// @phase D compact
// Phase D for one batch of listed faces: warp w owns the 8x4 block w, no CTA barrier inside.
__device__ void soft_batch(FwdSmem& s, const SoftGeom& G, SoftState& st)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned open32 = __ballot_sync(full, st.open && st.c < G.knum);
    if (!open32) return;
    const int lcount = s.lcount;
    unsigned int* ring_m = s.ring_m[warp];
    unsigned short* ring_li = s.ring_li[warp];
    int pn = 0;                                     // ring entries waiting
    for (int i0 = 0; i0 < lcount; i0 += 32) {
        const int li = i0 + lane;
        unsigned m32 = 0u;                          // open pixels of the block inside this lane's face's expanded bbox
        if (li < lcount) {
            const unsigned sm = s.smask[li];
            const unsigned cm = (sm >> G.bx) & 0xffu, rm = (sm >> (16 + G.by)) & 0xfu;
            m32 = (cm * 0x01010101u) & (((rm * 0x00204081u) & 0x01010101u) * 0xffu) & open32;
        }
        const unsigned bal = __ballot_sync(full, m32 != 0u);
        if (!bal) continue;
        if (m32) {
            const int pos = pn + __popc(bal & lt);
            ring_m[pos] = m32;
            ring_li[pos] = (unsigned short)li;
        }
        pn += __popc(bal);
        __syncwarp(full);
        if (pn >= 32) {
            soft_round(s, G, st, 32, open32);
            const int rest = pn - 32;
            unsigned mv_m = 0u;
            unsigned short mv_l = 0;
            if (lane < rest) { mv_m = ring_m[32 + lane]; mv_l = ring_li[32 + lane]; }
            __syncwarp(full);
            if (lane < rest) { ring_m[lane] = mv_m; ring_li[lane] = mv_l; }
            pn = rest;
            __syncwarp(full);
            if (!open32) return;                    // every pixel of the block has its K faces
        }
    }
    if (pn > 0) soft_round(s, G, st, pn, open32);
}

// @phase fill untouched
// One WARP fills a 16x16 tile of one [H,W,CH] image: 128-bit stores, 2*CH per lane
template <int CH>
__device__ __forceinline__ void fill_full_tile_warp(float* __restrict__ img, int width, int tx0, int ty0, float val)
{
    constexpr int RV = TILE * CH / 4;                    // float4 per tile row
    const int lane = threadIdx.x & 31;
    const float4 v = make_float4(val, val, val, val);
#pragma unroll
    for (int i0 = 0; i0 < TILE * RV; i0 += 32) {
        const int i = i0 + lane;
        const int r = i / RV, c = i - r * RV;
        reinterpret_cast<float4*>(img + ((size_t)(ty0 + r) * width + tx0) * CH)[c] = v;
    }
}

// one warp fills rows [0,th) x [0,tw) of one [H,W,ch] image tile with `val`
__device__ __forceinline__ void fill_tile_warp(float* __restrict__ img, int width, int ch, int tx0, int ty0, int tw, int th, float val)
{
    const bool aligned = (((size_t)width * ch) & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (tw == TILE && th == TILE && aligned && ch <= 4) {          // tx0 is a multiple of 16: rows start 16 B aligned
        switch (ch) {
            case 1: fill_full_tile_warp<1>(img, width, tx0, ty0, val); return;
            case 2: fill_full_tile_warp<2>(img, width, tx0, ty0, val); return;
            case 3: fill_full_tile_warp<3>(img, width, tx0, ty0, val); return;
            default: fill_full_tile_warp<4>(img, width, tx0, ty0, val); return;
        }
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
__device__ __noinline__ void fill_untouched_warp(const FwdParams& P, int packed_tile)
{
    int b, ty, tx;
    unpack_tile(packed_tile, b, ty, tx);
    const int tx0 = tx * TILE, ty0 = ty * TILE;
    const int tw = min(TILE, P.width - tx0), th = min(TILE, P.height - ty0);
    const size_t img_pix = (size_t)b * P.height * P.width;
    if ((threadIdx.x & 31) == 0 && P.min_group >= 0) atomicMin(P.out_min, f2ord(0.0f));
    for (int g = 0; g < P.n_out; g++) fill_tile_warp(P.out[g] + img_pix * P.out_ch[g], P.width, P.out_ch[g], tx0, ty0, tw, th, 0.0f);
    fill_tile_warp(P.improb + img_pix, P.width, 1, tx0, ty0, tw, th, 0.0f);
    fill_tile_warp(reinterpret_cast<float*>(P.imidx + img_pix), P.width, 1, tx0, ty0, tw, th, 0.0f);
    fill_tile_warp(P.imcomp + img_pix, P.width, 1, tx0, ty0, tw, th, 1.0f);
}

// @phase E write
// Phase E: the tile's copy of tensor `t` (plane of 256 * ch floats, row-major) -> global memory
__device__ __forceinline__ void write_plane(const float* __restrict__ plane, float* __restrict__ img, int width, int ch,
                                            int tx0, int ty0, int tw, int th)
{
    const int tid = threadIdx.x;
    const bool vec = (tw == TILE) && (th == TILE) && ((((size_t)width * ch) & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (vec) {
        const int rv = 4 * ch;                                   // float4 per tile row
        for (int v = tid; v < TILE * rv; v += NT) {
            const int r = v / rv, c = v - r * rv;
            reinterpret_cast<float4*>(img + ((size_t)(ty0 + r) * width + tx0) * ch)[c] = reinterpret_cast<const float4*>(plane)[v];
        }
    } else {
        const int rowf = tw * ch;
        for (int i = tid; i < th * rowf; i += NT) {
            const int r = i / rowf, x = i - r * rowf;
            img[((size_t)(ty0 + r) * width + tx0) * ch + x] = plane[r * TILE * ch + x];
        }
    }
}

// @phase prologue
__global__ void __launch_bounds__(NT, DIBR_FWD_CTAS_PER_SM)
dibr_forward_kernel(const __grid_constant__ FwdParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
    float* const stage = reinterpret_cast<float*>(smem_raw + sizeof(FwdSmem));     // (D + 3) planes of the tile's outputs
    const unsigned full_mask = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    const int ntiles = tiles_x * tiles_y * P.batch;
    const int D = P.num_attr;
    const float ex = P.expand_mul;
    // ---- the plan (set-up: plan_tiles_kernel): tiles by cost bucket.  bstart[l] = first position of bucket 31-l in
    //      heaviest-first order; bucket 0 (empty bitmaps) comes last.
    if (warp == 0) {
        static_assert(ORDER_BUCKETS == 32, "one bucket per lane");
        const int n = __ldg(P.order_cnt + (ORDER_BUCKETS - 1 - lane));
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full_mask, incl, o);
            if (lane >= o) incl += t;
        }
        s.bstart[lane] = incl - n;
        if (lane == 31) s.bstart[32] = incl;
    }
    if (tid == 0) mbar_init(&s.mbar, 1);
    __syncthreads();
    const int n_touched = s.bstart[31], n_empty = s.bstart[32] - s.bstart[31];
    const int n_items = n_touched + (n_empty + NWARP - 1) / NWARP;
    const int my_start = s.bstart[lane];
    // staging planes: output groups in order (channel d at chan_off[d] + pix * chan_stride[d]), then improb, imcomp, imidx
    float* const st_prob = stage + D * TILE * TILE;
    float* const st_comp = st_prob + TILE * TILE;
    int* const st_idx = reinterpret_cast<int*>(st_comp + TILE * TILE);
    int parity = 0;
    unsigned mbar_parity = 0u;
#ifdef DIBR_PHASE_TIMING
    long long t_phase = clock64();
#endif

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (item >= n_touched) {
            // untouched tiles: one warp each
            const int j = (item - n_touched) * NWARP + warp;
            if (j < n_empty) fill_untouched_warp(P, __ldg(P.order_seg + j));
            PHASE_MARK(6);
            continue;
        }
        int tile;
        {
            const unsigned le = __ballot_sync(full_mask, my_start <= item);        // lane 31 (bucket 0) starts at n_touched > item
            const int l = 31 - __clz(le);
            tile = __ldg(P.order_seg + (size_t)(ORDER_BUCKETS - 1 - l) * ntiles + (item - __shfl_sync(full_mask, my_start, l)));
        }
#ifdef DIBR_PHASE_TIMING
        if (tid == 0) atomicAdd(&g_phase[7], 1ull);
#endif
        int b, tile_y, tile_x;
        unpack_tile(tile, b, tile_y, tile_x);
        const int tile_in = tile_y * tiles_x + tile_x;
        const int tx0 = tile_x * TILE, ty0 = tile_y * TILE;
        TileGeom T;
        T.tw = min(TILE, P.width - tx0); T.th = min(TILE, P.height - ty0);
        const int tw = T.tw, th = T.th;
        const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
        const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
        const size_t img_pix = (size_t)b * P.height * P.width;
        const FaceRec* __restrict__ recs = P.recs + f_lo;

        // @phase tile setup
        // ---- tile set-up --------------------------------------------------------------------------------------
        if (tid < TILE) {
            s.xs[tid] = (tid < tw) ? __ldg(P.xs + tx0 + tid) : 3.0e38f;
        } else if (tid < 2 * TILE) {
            const int r = tid - TILE;
            s.ys[r] = (r < th) ? __ldg(P.ys + ty0 + r) : -3.0e38f;
        }
        s.u.cov.zkey[tid] = 0ull;
        T.inv_dx = 0.5f * (float)P.width / (float)P.multiplier;     // pixel pitch is 2m/W
        T.inv_dy = 0.5f * (float)P.height / (float)P.multiplier;
        {
            const int w0 = f_lo >> 5;
            T.nw = ((f_hi - 1) >> 5) - w0 + 1;
            T.id0 = (w0 << 5) - f_lo;
            T.words = P.bins + (size_t)tiles_x * tiles_y * ((size_t)w0 + b) + (size_t)tile_in * T.nw;
        }
        // (the first barrier of expand_batch orders these writes before their readers)

        // @phase batch loop
        // ---- phases A + B, one batch of LCAP listed faces at a time -----------------------------------------------
        int nbatch = 0;
        {
            int wpos = 0, nprev = 0;
            do {
                wpos = expand_batch(s, wpos, T, parity);
                gather_records(s, recs, mbar_parity);
                PHASE_MARK(0);
                prep_faces(s, T, ex, true);
                PHASE_MARK(1);
                const int lcount = s.lcount;
                if (s.rcount > 0) coverage(s, nprev);
                if (wpos < T.nw || nbatch > 0) {
                    // several batches: pin down this batch's winners while its id list is still in shared memory
                    const unsigned long long key = s.u.cov.zkey[tid];
                    if (key != 0ull) {
                        const unsigned rank = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
                        if (rank >= (unsigned)nprev) s.u.cov.wf[tid] = (unsigned)s.lid[rank - nprev];
                    }
                    __syncthreads();
                }
                PHASE_MARK(2);
                nprev += lcount;
                nbatch++;
            } while (wpos < T.nw);
        }
        const bool single = (nbatch == 1);

        // @phase C resolve
        // ---- phase C: resolve, one pixel per thread (a warp covers two tile rows) --------------------------------
        const float* __restrict__ fattr = P.face_attr + (size_t)f_lo * 3 * D;
        bool unc = false;
        {
            const int lx = tid & (TILE - 1), ly = tid >> 4;
            const bool val = (lx < tw) && (ly < th);
            int fw = -1;
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 c1 = make_float2(0.f, 0.f);
            if (val) {
                const unsigned long long key = s.u.cov.zkey[tid];
                if (key != 0ull) {
                    if (single) {
                        const int li = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
                        fw = s.lid[li];
                        c0 = s.rec[li * REC_V4];
                        const float4 t1 = s.rec[li * REC_V4 + 1];
                        c1 = make_float2(t1.x, t1.y);
                    } else {
                        fw = (int)s.u.cov.wf[tid];
                        const float4* rp = reinterpret_cast<const float4*>(recs + fw);
                        c0 = __ldg(rp);
                        const float4 t1 = __ldg(rp + 1);
                        c1 = make_float2(t1.x, t1.y);
                    }
                }
            }
            // winners go on the backward's colour work list (run-length de-duplicated along the row)
            const int prev = __shfl_up_sync(full_mask, fw, 1);
            const bool lead = fw >= 0 && (lx == 0 || prev != fw);
            const unsigned fl = lead ? __ldcg(&P.face_flags[f_lo + fw]) : 1u;
            float vmin = 3.0e38f;                        // minimum of output group P.min_group
            float v[DIBR_MAX_ATTR_INTERNAL];
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) v[d] = 0.f;
            if (fw >= 0) {
                FaceRec r;
                r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
                r.az = r.bz = r.cz = 0.f;
                const FaceK fk = make_facek(r);
                float w0, w1, w2;
                bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2);
                const float* a = fattr + (size_t)fw * 3 * D;
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
            // the tile's copy of the outputs (pixels outside the image are never written out)
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) {
                if (d < D) {
                    stage[P.chan_off[d] + tid * P.chan_stride[d]] = v[d];
                    if (val && ((P.min_mask >> d) & 1u)) vmin = fminf(vmin, v[d]);
                }
            }
            st_prob[tid] = (fw >= 0) ? 1.0f : 0.0f;      // uncovered: empty product; phase D overwrites the pixels it reaches
            st_comp[tid] = (fw >= 0) ? 0.0f : 1.0f;
            st_idx[tid] = fw + 1;                        // uncovered: 0, may become -(K-th face + 1) in phase D
            s.cnt[tid] = (fw >= 0) ? 255 : 0;
            unc = val && fw < 0;
            // append first-time winners to the colour list: one counter atomic per warp
            {
                bool isnew = false;
                const int g = f_lo + max(fw, 0);
                if (lead && (fl & 1u) == 0u) isnew = (atomicOr(&P.face_flags[g], 1u) & 1u) == 0u;
                const unsigned nb = __ballot_sync(full_mask, isnew);
                if (nb) {
                    const int leader = __ffs(nb) - 1;
                    int lb = 0;
                    if (lane == leader) lb = atomicAdd(&P.list_counts[0], __popc(nb));
                    lb = __shfl_sync(full_mask, lb, leader);
                    if (isnew) P.color_list[lb + __popc(nb & ((1u << lane) - 1u))] = g;
                }
            }
            if (P.min_group >= 0) {
                const unsigned ov = __reduce_min_sync(full_mask, f2ord(vmin));
                if (lane == 0 && ov != f2ord(3.0e38f)) atomicMin(P.out_min, ov);
            }
        }
        const int tile_unc = __syncthreads_or(unc ? 1 : 0);        // also: staging + cnt[] complete, the z-buffer is dead
        PHASE_MARK(3);

        // @phase D outer
        // ---- phase D: soft silhouette.  Thread = pixel, block-major: warp w owns the 8x4 block w ------------------
        if (tile_unc && P.knum > 0) {
            SoftGeom G;
            G.zscale = (float)P.delta / ((float)P.multiplier * (float)P.multiplier);
            G.sentinel = 4.0f * (float)P.multiplier * (float)P.multiplier;
            G.knum = P.knum;
            G.bx = (warp % NBX) * BW; G.by = (warp / NBX) * BH;
            const int lx = G.bx + (lane & 7), ly = G.by + (lane >> 3);
            const int pix = ly * TILE + lx;
            SoftState st;
            st.q = 0.f; st.cc = 1.f; st.c = 0; st.kth = -1;
            st.open = (lx < tw) && (ly < th) && s.cnt[pix] == 0;
            int wpos = 0;
            for (;;) {
                if (!single) {
                    wpos = expand_batch(s, wpos, T, parity);
                    gather_records(s, recs, mbar_parity);
                    prep_faces(s, T, ex, false);
                }
                soft_batch(s, G, st);
                __syncthreads();                                    // soft_used complete
                // hand the faces that contributed to the backward's work list
                const int lcount = s.lcount;
                for (int li0 = 0; li0 < lcount; li0 += NT) {
                    const int li = li0 + tid;
                    const bool used = (li < lcount) && s.soft_used[li];
                    const int g = used ? f_lo + s.lid[li] : 0;
                    bool isnew = false;
                    if (used && (__ldcg(&P.face_flags[g]) & 2u) == 0u) isnew = (atomicOr(&P.face_flags[g], 2u) & 2u) == 0u;
                    const unsigned bal = __ballot_sync(full_mask, isnew);
                    if (bal) {
                        const int leader = __ffs(bal) - 1;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(&P.list_counts[1], __popc(bal));
                        base = __shfl_sync(full_mask, base, leader);
                        if (isnew) P.soft_list[base + __popc(bal & ((1u << lane) - 1u))] = g;
                    }
                }
                if (single || wpos >= T.nw) break;
                // stop early once every uncovered pixel has its K faces (also: everybody is done with this batch's lists)
                if (!__syncthreads_or((st.open && st.c < G.knum) ? 1 : 0)) break;
            }
            if (st.open) {
                st_prob[pix] = fminf(st.q, 1.0f);                   // the recurrence can overshoot 1 by an ulp
                st_comp[pix] = st.cc;
                if (st.kth >= 0) st_idx[pix] = -(st.kth + 1);
            }
            __syncthreads();
        }
        PHASE_MARK(4);

        // @phase E write
        // ---- phase E: the tile leaves shared memory, every tensor written once with 128-bit stores ------------------
        {
            int off = 0;
            for (int g = 0; g < P.n_out; g++) {
                const int ch = P.out_ch[g];
                write_plane(stage + off, P.out[g] + img_pix * ch, P.width, ch, tx0, ty0, tw, th);
                off += ch * TILE * TILE;
            }
            write_plane(st_prob, P.improb + img_pix, P.width, 1, tx0, ty0, tw, th);
            write_plane(st_comp, P.imcomp + img_pix, P.width, 1, tx0, ty0, tw, th);
            write_plane(reinterpret_cast<const float*>(st_idx), reinterpret_cast<float*>(P.imidx + img_pix), P.width, 1, tx0, ty0, tw, th);
        }
        __syncthreads();                                            // the next tile reuses everything
        PHASE_MARK(5);
    }
}

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

#ifdef DIBR_PHASE_TIMING
extern "C" void dibr_debug_phase_cycles(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_phase, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof(z)); }
}
#endif

// forward implementation switch for the A/B (profiles/): DIBR_FWD_IMPL=2 in the environment selects the previous design
static int forward_impl() {
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("DIBR_FWD_IMPL");
        impl = (e && e[0] == '2') ? 2 : 3;
    }
    return impl;
}

int launch_forward(const FwdParams& P, cudaStream_t stream)
{
    if (forward_impl() == 2) return launch_forward_v2(P, stream);
    const size_t smem = sizeof(FwdSmem) + sizeof(float) * (size_t)(P.num_attr + 3) * TILE * TILE;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    // per-device launch configuration, computed once
    static int grid_of[64];
    static size_t smem_set[64];
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (smem_set[dev] < smem) {
        e = cudaFuncSetAttribute(dibr_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FwdSmem) + sizeof(float) * (DIBR_MAX_ATTR_INTERNAL + 3) * TILE * TILE));
        if (e != cudaSuccess) return (int)e;
        smem_set[dev] = sizeof(FwdSmem) + sizeof(float) * (DIBR_MAX_ATTR_INTERNAL + 3) * TILE * TILE;
        int sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        grid_of[dev] = sms * DIBR_FWD_CTAS_PER_SM;
    }
    const int ntiles = ((P.width + TILE - 1) / TILE) * ((P.height + TILE - 1) / TILE) * P.batch;
    const int grid = min(grid_of[dev], ntiles);
    dibr_forward_kernel<<<grid, NT, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
