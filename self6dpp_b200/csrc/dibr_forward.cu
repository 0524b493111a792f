// Forward pass of the B200 DIB-R rasterizer: one CTA (256 threads, one per pixel) per 16x16 screen tile, 47.5 KB of
// shared memory and <= 64 registers so four CTAs share an SM (194 KB of the SM's 256 KB: the 196 KB carve-out leaves
// 60 KB of L1 -- at 53 KB per CTA the next carve-out step took half of that L1 away and cost 5 %).  CTA i takes the
// i-th tile of the plan the set-up call left (plan_tiles_kernel: tiles bucketed by list length, heaviest first); tiles
// whose bitmap is empty are filled by one warp each.  Everything after the list build runs out of shared memory; the
// only trips to L2 are the tile's bitmap, ONE gather of the listed faces' records, the winners' attributes and the
// work-list flags.
//
//   phase A  read the tile's face bitmap (one bit per face of the image, set by the set-up kernel's binning), scan the
//            popcounts across the CTA and expand the set bits into the ascending list of face ids -- no per-tile scan
//            over all faces and no sort.  Then the records of the listed faces are gathered into shared memory
//            (corners, depths) and the front faces that really hold a pixel centre of the tile go on the raster list.
//   phase B  face-parallel coverage with 8 lanes per face: barycentric solve in the frozen fp32 order and a 64-bit
//            shared-memory atomicMax on (orderable z | ~rank in the list).  The winner is the face with the largest z
//            and, on ties, the smallest index -- what the reference's ascending loop with a strict '>' produces,
//            independent of traversal order.
//   phase C  resolve: per pixel recompute the winner's weights, interpolate the D attributes and write every
//            output tensor / improb=1 / imidx.
//   phase D  soft silhouette for uncovered pixels: (1) collect, one 8x4 pixel block per warp: every lane turns one
//            listed face into a 32-bit mask of the block's pixels inside its expanded bbox, the masks are dealt back
//            to the pixel lanes in ascending face order (first K per pixel); (2) all (pixel, face) pairs of the TILE
//            are split evenly over the 256 threads and evaluated (distance, exp) in place; (3) each pixel folds its
//            own results in face order.
//
// Replaces kaolin v0.1's dr_cuda_forward_render_batch + dr_cuda_forward_prob_batch, which the reference calls at
// lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 (one thread per pixel looping over ALL faces,
// SURVEY.md 8(a) rows a6/a7).  Instead of the reference's five B x H x W x 30 scratch tensors (rasterizer.py:144-148)
// only the K-th accepted face id is kept, folded into imidx (see include/dibr_b200.h).
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

#define PHASE_MARK(k) do { } while (0)

constexpr int NWARP = FWD_THREADS / 32;
#ifndef DIBR_HITCAP
#define DIBR_HITCAP 30
#endif
constexpr int HITCAP = DIBR_HITCAP;     // collected faces per pixel per pass of phase D (pixels with more take another pass)
constexpr int BW = 8, BH = 4;           // pixel block of one warp in phase D
constexpr int NBX = TILE / BW;
#ifndef DIBR_RASTER_LANES
#define DIBR_RASTER_LANES 8
#endif
constexpr int RASTER_LANES = DIBR_RASTER_LANES;      // lanes per face in phase B
static_assert(NBX * (TILE / BH) == NWARP, "one 8x4 block per warp");
static_assert(LCAP >= 32, "one bitmap word must fit an empty list");
static_assert(LCAP <= 512 && LCAP % 16 == 0 && TILE == 16, "rlist packing: 9-bit list index, 4-bit pixel coordinates");

struct FwdSmem {
    float4 c0[LCAP];                            //  8 KB  ax ay bx by   (x multiplier)
    float2 c1[LCAP];                            //  4 KB  cx cy
    int lid[LCAP];                              //  2 KB  local face ids, ascending
    union {                                     // 30 KB
        struct {
            unsigned long long zkey[TILE * TILE];   //    z-buffer (phases B, C)
            unsigned int rlist[LCAP];           //        raster candidates (phase B)
            float z[3][LCAP];                   //        view-space depth of the corners (coverage only)
            unsigned int wf[TILE * TILE];       //        winner's face id, committed per batch when a tile needs several
            int big[BIGCAP];
            unsigned int cq[FWD_THREADS / 32][64];  //    per-warp queue of (face, pixel) pairs that passed the cheap inside test (phase B)
            int4 vid[LCAP];                     //        fused mode: attribute rows of the raster candidates' corners (fetched with the records,
                                                //        so the resolve does not chain two gathers)
        } ab;
        unsigned int hits[HITCAP][FWD_THREADS]; //        phase D: list index of the k-th face of a pixel, then its result
    } u;
    unsigned short E[FWD_THREADS];              // phase D: exclusive scan of the pixels' hit counts inside their warp
    unsigned char nhc[FWD_THREADS];             // phase D: hit count of each pixel in the current pass
    unsigned char cnt[TILE * TILE];             // accepted faces per pixel (255 = covered)
    unsigned char soft_used[LCAP];              // listed faces that entered some pixel's soft product
    unsigned int smask[LCAP];                   //  2 KB  phase D: tile columns (bits 0-15) / rows (16-31) inside the face's expanded bbox
    float xs[TILE], ys[TILE];
    int warp_tot[2][NWARP];
    int nbig, lcount, rcount, rnext;            // rnext: next unclaimed entry of the raster list (warps draw four faces at a time)
};

// Faces that did `bit`-type work (1: won a pixel, 2: entered a soft product) are flagged with a fire-and-forget atomic (RED):
// the backward compacts the flags into its two work lists before it starts (prepare_backward_kernel), so no tile waits
// for a returning atomic on a list counter that every CTA of the grid is hammering.
__device__ __forceinline__ void mark_face(const FwdParams& P, int g, unsigned bit) {
    atomicOr(&P.face_flags[g], bit);
}

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

// first column c in [0,n] with xs[c] >= x (xs ascending, pitch 1/inv_dx): arithmetic guess + exact fix-up
__device__ __forceinline__ int col_first_ge(const float* xs, int n, float x, float inv_dx) {
    int c = (int)fminf(fmaxf(ceilf((x - xs[0]) * inv_dx), 0.f), (float)n);
    while (c > 0 && xs[c - 1] >= x) c--;
    while (c < n && xs[c] < x) c++;
    return c;
}
// first row r in [0,n] with ys[r] < y (ys descending)
__device__ __forceinline__ int row_first_lt(const float* ys, int n, float y, float inv_dy) {
    int r = (int)fminf(fmaxf(floorf((ys[0] - y) * inv_dy) + 1.0f, 0.f), (float)n);
    while (r > 0 && ys[r - 1] < y) r--;
    while (r < n && ys[r] >= y) r++;
    return r;
}

struct TileGeom {
    int tw, th, tx0, ty0;
    float inv_dx, inv_dy;
    const uint32_t* words;          // this tile's face bitmap (set-up kernel: bin_face)
    int nw;                         // its length in 32-face words
    int id0;                        // local face id of bit 0 of word 0  (<= 0)
};

// Phase A.  Expands the tile's face bitmap from word `wpos` on into the ascending list of local face ids, until the
// list is full (a batch always ends on a word boundary); gathers the listed faces' records; when `raster` also builds
// the raster list.  Returns the first unread word.  Uniform across the CTA.
template <bool FUSED>
__device__ int fill_list(FwdSmem& s, const FwdParams& P, int f_lo, int wpos, const TileGeom& T, bool raster, int& parity)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int lcount = 0;
    if (tid == 0) { s.rcount = 0; s.rnext = 0; }         // ordered before the gather's atomicAdds by the barriers below; its last readers are past raster_list's barrier
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
        int taken = FWD_THREADS;                                // words of this round that go on the list
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
        if (taken < FWD_THREADS) {
            // the thread of the last taken word knows the new length
            if (taken == 0) break;
            if (tid == taken - 1) s.lcount = gi;
            __syncthreads();
            lcount = s.lcount;
            wpos += taken;
            break;
        }
        lcount += total;
        wpos += FWD_THREADS;
    }
    wpos = min(wpos, T.nw);
    if (tid == 0) s.lcount = lcount;    // uniform across the CTA: published for the phases after this call
    __syncthreads();
    // ---- gather the listed faces' records (the one L2 round trip of the list) and build the raster list:
    //      front faces with a non-empty pixel range (packed: list index | c0 | nc-1 | r0 | nr-1), any order
    const FaceRec* __restrict__ recs = P.recs + f_lo;
    for (int i0 = 0; i0 < lcount; i0 += FWD_THREADS) {
        const int i = i0 + tid;
        bool keep = false;
        unsigned int packed = 0u;
        if (i < lcount) {
            const float4* rp = reinterpret_cast<const float4*>(recs + s.lid[i]);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
            // pixel ranges of the bbox and of the expanded bbox (exact, from the set-up kernel), clipped to the tile
            const uint4 rg = __ldg(reinterpret_cast<const uint4*>(rp + 3));
            s.c0[i] = r0;
            s.c1[i] = make_float2(r1.x, r1.y);
            {
                const int e0 = min(max((int)(rg.z & 0xffffu) - T.tx0, 0), T.tw), e1 = min(max((int)(rg.z >> 16) - T.tx0, 0), T.tw);
                const int q0 = min(max((int)(rg.w & 0xffffu) - T.ty0, 0), T.th), q1 = min(max((int)(rg.w >> 16) - T.ty0, 0), T.th);
                s.smask[i] = (e1 > e0 && q1 > q0) ? (((1u << e1) - (1u << e0)) | (((1u << q1) - (1u << q0)) << 16)) : 0u;
            }
            if (raster) { s.u.ab.z[0][i] = r1.z; s.u.ab.z[1][i] = r1.w; s.u.ab.z[2][i] = r2.x; }
            if (raster && r2.y >= 0.0f) {                       // front face (K1 culls normalz < 0)
                const int c0 = min(max((int)(rg.x & 0xffffu) - T.tx0, 0), T.tw), c1 = min(max((int)(rg.x >> 16) - T.tx0, 0), T.tw);
                const int q0 = min(max((int)(rg.y & 0xffffu) - T.ty0, 0), T.th), q1 = min(max((int)(rg.y >> 16) - T.ty0, 0), T.th);
                if (c1 > c0 && q1 > q0) {
                    keep = true;
                    if (FUSED) s.u.ab.vid[i] = __ldg(P.va.fvid + f_lo + s.lid[i]);
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
                if (keep) s.u.ab.rlist[rb + __popc(bal & lt)] = packed;
            }
        }
    }
    __syncthreads();
    return wpos;
}

__constant__ unsigned c_inv16[17] = {0u, 65537u, 32769u, 21846u, 16385u, 13108u, 10923u, 9363u, 8193u, 7282u, 6554u, 5958u, 5462u, 5042u, 4682u, 4370u, 4097u};

struct RasterEntry { int li, c0, nc, r0, nr; };
__device__ __forceinline__ RasterEntry unpack_entry(unsigned int p) {
    RasterEntry e;
    e.li = p & 511; e.c0 = (p >> 9) & 15; e.nc = ((p >> 13) & 15) + 1; e.r0 = (p >> 17) & 15; e.nr = ((p >> 21) & 15) + 1;
    return e;
}

__device__ __forceinline__ FaceK facek_from_list(const FwdSmem& s, int li) {
    const float4 a = s.c0[li];
    const float2 b = s.c1[li];
    FaceRec r;
    r.ax = a.x; r.ay = a.y; r.bx = a.z; r.by = a.w; r.cx = b.x; r.cy = b.y;
    r.az = s.u.ab.z[0][li]; r.bz = s.u.ab.z[1][li]; r.cz = s.u.ab.z[2][li];
    return make_facek(r);
}

__device__ __forceinline__ void raster_pixel(FwdSmem& s, const FaceK& fk, unsigned rank, int lx, int ly) {
    float w0, w1, w2;
    if (!bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2)) return;
    float z0 = blend(w0, w1, w2, fk.az, fk.bz, fk.cz);
    if (!(z0 > -1000.0f)) return;                 // "z0 <= znow" against the initial depth -1000
    z0 = z0 + 0.0f;                               // -0 -> +0 so equal depths compare equal
    const unsigned long long key = ((unsigned long long)f2ord(z0) << 32) | (unsigned long long)(0xffffffffu - rank);
    atomicMax(&s.u.ab.zkey[ly * TILE + lx], key);
}

// Phase B.  `nprev`: faces listed by earlier batches of this tile (ranks keep ascending across batches).
//
// Four in five pixel centres of a face's bbox lie outside the face.  They are turned away by a CHEAP conservative test --
// the signs of the two edge functions and of their sum against k3, no division, with margins far above the rounding of
// the exact solve -- and only the survivors, queued per warp and taken 32 at a time with every lane busy, get the exact
// barycentric solve in the frozen fp32 order (two IEEE divisions) that decides coverage and depth.  A pixel the cheap
// test rejects is one the exact test rejects too (proof in the comments below), so the face ids stay bit-exact.
__device__ __forceinline__ void raster_queued(FwdSmem& s, unsigned entry, int nprev) {
    const int li = (int)(entry & 511u);
    const FaceK fk = facek_from_list(s, li);
    raster_pixel(s, fk, (unsigned)(nprev + li), (int)((entry >> 9) & 15u), (int)(entry >> 13));
}

__device__ void raster_list(FwdSmem& s, int nprev)
{
    const unsigned full_mask = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rcount = s.rcount;
    const unsigned lt = (1u << lane) - 1u;
    unsigned int* const cq = s.u.ab.cq[warp];
    int qn = 0;                                                         // queued candidates of this warp (< 32 between turns)
    // ---- RASTER_LANES lanes per face; faces with many pixels in the tile are deferred to the whole CTA
    const int ql = tid % RASTER_LANES;
    for (;;) {
        int e0 = 0;
        if (lane == 0) e0 = atomicAdd(&s.rnext, 32 / RASTER_LANES);     // the faces' pixel counts differ: warps draw work instead of striding (-1.5 us)
        e0 = __shfl_sync(full_mask, e0, 0);
        if (e0 >= rcount) break;
        const int e = e0 + lane / RASTER_LANES;
        int npx = 0, li = 0, c0 = 0, r0 = 0, nc = 1;
        float ax = 0.f, ay = 0.f, em = 0.f, ep = 0.f, en_ = 0.f, eq = 0.f, sg = 1.f, lim = 0.f;
        bool cheap_ok = false;
        if (e < rcount) {
            const unsigned int packed = s.u.ab.rlist[e];
            const RasterEntry en = unpack_entry(packed);
            npx = en.nc * en.nr;
            if (npx > BIG_AREA) {
                if (ql == 0) {
                    const int slot = atomicAdd(&s.nbig, 1);
                    if (slot < BIGCAP) s.u.ab.big[slot] = (int)packed;     // beyond BIGCAP: picked up by the rescan below
                }
                npx = 0;
            } else {
                li = en.li; c0 = en.c0; r0 = en.r0; nc = en.nc;
                const float4 a = s.c0[li];
                const float2 b = s.c1[li];
                ax = a.x; ay = a.y;
                em = __fsub_rn(a.z, a.x); ep = __fsub_rn(a.w, a.y);         // make_facek's m, p, n, q, k3
                en_ = __fsub_rn(b.x, a.x); eq = __fsub_rn(b.y, a.y);
                const float k3 = __fmaf_rn(em, eq, -__fmul_rn(en_, ep));
                // the margins below hold while the quotients k / k3 of the exact solve are ordinary fp32 divisions (|k3| >= 32,
                // div_eps) that can neither overflow nor flush to zero
                cheap_ok = fabsf(k3) >= 32.0f && fabsf(k3) <= 1.0e15f;
                sg = copysignf(1.0f, k3);
                lim = fabsf(k3) * 1.0001f;
            }
        }
        const int maxn = __reduce_max_sync(full_mask, npx);
        const unsigned inv = c_inv16[nc];                                  // 65536 / nc + 1: exact i / nc for i < 256
        for (int i0 = 0; i0 < maxn; i0 += RASTER_LANES) {
            const int i = i0 + ql;
            bool cand = false;
            unsigned entry = 0u;
            if (i < npx) {
                const int row = (int)(((unsigned)i * inv) >> 16);
                const int lx = c0 + (i - row * nc), ly = r0 + row;
                const float sx = __fsub_rn(s.xs[lx], ax), ty = __fsub_rn(s.ys[ly], ay);
                const float k1 = __fmaf_rn(sx, eq, -__fmul_rn(en_, ty)) * sg;      // w1 = k1 / k3, w2 = k2 / k3 (bary)
                const float k2 = __fmaf_rn(em, ty, -__fmul_rn(sx, ep)) * sg;
                // k1 sg < -1e-12: w1 = -|k1| / |k3| <= -1e-27 is an ordinary negative number -> bary() says outside.  Same for
                // w2.  Otherwise both are >= -3e-14, and (k1 + k2) sg > 1.0001 |k3| puts w1 + w2 above 1 by 1e-4, four hundred
                // times the rounding of w0 = (1 - w1) - w2 -> w0 < 0.  NaNs fail every comparison and stay candidates.
                cand = !cheap_ok || !(k1 < -1.0e-12f || k2 < -1.0e-12f || (k1 + k2) > lim);
                entry = (unsigned)li | ((unsigned)lx << 9) | ((unsigned)ly << 13);
            }
            const unsigned bal = __ballot_sync(full_mask, cand);
            if (cand) cq[qn + __popc(bal & lt)] = entry;
            qn += __popc(bal);
            if (qn >= 32) {
                __syncwarp();
                raster_queued(s, cq[qn - 32 + lane], nprev);
                qn -= 32;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if (lane < qn) raster_queued(s, cq[lane], nprev);
    __syncthreads();
    // ---- large faces: one pixel per thread
    const int nbig_all = s.nbig;
    if (nbig_all > 0) {
        auto whole = [&](unsigned int packed) {
            const RasterEntry en = unpack_entry(packed);
            const FaceK fk = facek_from_list(s, en.li);
            if (tid < en.nc * en.nr) raster_pixel(s, fk, (unsigned)(nprev + en.li), en.c0 + tid % en.nc, en.r0 + tid / en.nc);
        };
        for (int j = 0; j < min(nbig_all, BIGCAP); j++) whole((unsigned)s.u.ab.big[j]);
        if (nbig_all > BIGCAP) {
            for (int e = 0; e < rcount; e++) {
                const unsigned int packed = s.u.ab.rlist[e];
                const RasterEntry en = unpack_entry(packed);
                if (en.nc * en.nr <= BIG_AREA) continue;
                bool listed = false;
                for (int j = 0; j < BIGCAP; j++) listed |= ((unsigned)s.u.ab.big[j] == packed);
                if (!listed) whole(packed);
            }
        }
        __syncthreads();
        if (tid == 0) s.nbig = 0;           // everybody read it before the barrier; the next user sits behind fill_list's barriers
    }
}

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

// nothing near this tile: zeros everywhere.  One warp per tile.
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
    // imcomp is NOT written here: the backward reads it only at uncovered pixels inside some face's expanded range, i.e. on touched tiles
}

#ifndef DIBR_FWD_MIN_CTAS
#define DIBR_FWD_MIN_CTAS (1024 / DIBR_FWD_THREADS)
#endif
template <bool FUSED>      // FUSED: corner attributes through P.va (pass set up by dibr_setup_meshes), else from P.face_attr
__global__ void __launch_bounds__(FWD_THREADS, DIBR_FWD_MIN_CTAS)
dibr_forward_kernel(const __grid_constant__ FwdParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
    const unsigned full_mask = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- which tile: the plan (set-up: plan_tiles_kernel) lists the tiles by cost bucket, heaviest first.  Lane k
    //      looks at bucket 31-k: one load each, a warp scan finds the bucket that holds position blockIdx.x.
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    const int ntiles = tiles_x * tiles_y * P.batch;
    int tile;
    int4 desc;               // plan entry: tile id, the image's face range, offset of the tile's bitmap
    {
        // the plan's summary (written by the set-up CTA that planned the last image): lane l reads the first position of
        // bucket 31 - l, every lane the size of bucket 0 (untouched tiles) -- two independent loads, ONE trip to L2 before the
        // tile's plan entry can be fetched.  Surplus CTAs of the grid leave here.
        static_assert(ORDER_BUCKETS == 32, "one bucket per lane");
        const int start = __ldg(P.order_cnt + PLAN_START + lane);                      // first position of bucket 31 - lane
        const int n_untouched = __ldg(P.order_cnt);
        const int touched = __shfl_sync(full_mask, start, 31);                        // bucket 0 comes last
        if ((int)blockIdx.x >= touched + ((n_untouched + NWARP - 1) / NWARP)) return;
        if ((int)blockIdx.x >= touched) {
            // bucket 0 (empty bitmaps): one warp per tile, 8 tiles per CTA
            const int j = ((int)blockIdx.x - touched) * NWARP + warp;
#ifndef DIBR_EXP_NOFILL
            if (j < n_untouched) fill_untouched_warp(P, __ldg(P.order_seg + j));
#endif
            return;
        }
        const unsigned le = __ballot_sync(full_mask, start <= (int)blockIdx.x);
        const int src = 31 - __clz(le);
        desc = __ldg(P.order_desc + (size_t)(ORDER_BUCKETS - 1 - src) * ntiles + ((int)blockIdx.x - __shfl_sync(full_mask, start, src)));
        tile = desc.x;
    }
#ifdef DIBR_EXP_ONLYFILL
    return;
#endif
    int b, tile_y, tile_x;
    unpack_tile(tile, b, tile_y, tile_x);
    const int tx0 = tile_x * TILE, ty0 = tile_y * TILE;
    TileGeom T;
    T.tw = min(TILE, P.width - tx0); T.th = min(TILE, P.height - ty0); T.tx0 = tx0; T.ty0 = ty0;
    const int tw = T.tw, th = T.th;
    const int f_lo = desc.y, f_hi = desc.z;
    const int D = P.num_attr;
    const size_t img_pix = (size_t)b * P.height * P.width;
    float* __restrict__ improb = P.improb + img_pix;
    float* __restrict__ imcomp = P.imcomp + img_pix;
    int* __restrict__ imidx = P.imidx + img_pix;


    // ---- tile set-up ----------------------------------------------------------------------------
    if (tid < TILE) {
        s.xs[tid] = (tid < tw) ? __ldg(P.xs + tx0 + tid) : 3.0e38f;
    } else if (tid < 2 * TILE) {
        const int r = tid - TILE;
        s.ys[r] = (r < th) ? __ldg(P.ys + ty0 + r) : -3.0e38f;
    }
    if (tid == 0) { s.nbig = 0; s.lcount = 0; s.rcount = 0; }
    s.u.ab.zkey[tid] = 0ull;
    s.cnt[tid] = 0;
    T.inv_dx = 0.5f * (float)P.width / (float)P.multiplier;     // pixel pitch is 2m/W
    T.inv_dy = 0.5f * (float)P.height / (float)P.multiplier;
    {
        const int w0 = f_lo >> 5;
        T.nw = ((f_hi - 1) >> 5) - w0 + 1;
        T.id0 = (w0 << 5) - f_lo;
        T.words = P.bins + (size_t)(unsigned)desc.w;
    }
    // no barrier here: nothing reads these before fill_list's first barrier

    int nbatch = 0, parity = 0;
    {
        int wpos = 0, nprev = 0;
        PHASE_MARK(0);
        while (wpos < T.nw) {
            wpos = fill_list<FUSED>(s, P, f_lo, wpos, T, true, parity);
            PHASE_MARK(1);
            const int lcount = s.lcount;
#ifndef DIBR_EXP_NORASTER
            if (s.rcount > 0) raster_list(s, nprev);
#endif
            if (wpos < T.nw || nbatch > 0) {
                // several batches: pin down this batch's winners while its id list is still in shared memory
                const unsigned long long key = s.u.ab.zkey[tid];
                if (key != 0ull) {
                    const unsigned rank = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
                    if (rank >= (unsigned)nprev) s.u.ab.wf[tid] = (unsigned)s.lid[rank - nprev];
                }
                __syncthreads();
            }
            PHASE_MARK(2);
            nprev += lcount;
            nbatch++;
        }
    }
    const bool single = (nbatch == 1);

    // ---- phase C: resolve, one pixel per thread (a warp covers two tile rows) --------------------------------
    const float* __restrict__ fattr = P.face_attr + (size_t)f_lo * 3 * D;
    const FaceRec* __restrict__ recs = P.recs + f_lo;
    bool unc = false;
    {
        const int lx = tid & (TILE - 1), ly = tid >> 4;
        const bool val = (lx < tw) && (ly < th);
        int fw = -1;
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 c1 = make_float2(0.f, 0.f);
        int wli = 0;                                    // fused mode: the winner's list index (single batch): its corner depths
        if (val) {
            const unsigned long long key = s.u.ab.zkey[ly * TILE + lx];
            if (key != 0ull) {
                if (single) {
                    const int li = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
                    fw = s.lid[li];
                    c0 = s.c0[li]; c1 = s.c1[li];
                    wli = li;
                } else {
                    fw = (int)s.u.ab.wf[ly * TILE + lx];
                    const float4* rp = reinterpret_cast<const float4*>(recs + fw);
                    c0 = __ldg(rp);
                    const float4 t1 = __ldg(rp + 1);
                    c1 = make_float2(t1.x, t1.y);
                }
            }
        }
        // winners are flagged for the backward's colour work list (run-length de-duplicated along the row)
        const int prev = __shfl_up_sync(full_mask, fw, 1);
        const bool lead = fw >= 0 && (lx == 0 || prev != fw);
        if (lead) mark_face(P, f_lo + fw, 1u);
        float vmin = 3.0e38f;                        // minimum of output group P.min_group
        if (val) {
            const size_t gp = (size_t)(ty0 + ly) * P.width + (tx0 + lx);
            const size_t px = img_pix + gp;
            float v[DIBR_MAX_ATTR_INTERNAL];
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) v[d] = 0.f;
            float vone = 0.f, vdep = 0.f;         // fused mode: the interpolated ones / depth channels
            if (fw >= 0) {
                FaceRec r;
                r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
                r.az = r.bz = r.cz = 0.f;
                const FaceK fk = make_facek(r);
                float w0, w1, w2;
                bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2);
                if (FUSED) {
                    // corner attributes = [row of the vertex table | 1 | -view z] through the face's three row ids
                    const int4 id = single ? s.u.ab.vid[wli] : __ldg(P.va.fvid + f_lo + fw);
                    const int A = P.va.dim, ones = P.va.flags & 1, dep = (P.va.flags >> 1) & 1;
                    const float* a0 = P.va.table + (size_t)id.x * P.va.stride;
                    const float* a1 = P.va.table + (size_t)id.y * P.va.stride;
                    const float* a2 = P.va.table + (size_t)id.z * P.va.stride;
                    if (P.va.vec) {                      // rows padded to a multiple of 16 B
#pragma unroll
                        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d += 4) {
                            if (d < A) {
                                const float4 r0 = __ldg(reinterpret_cast<const float4*>(a0 + d));
                                const float4 r1 = __ldg(reinterpret_cast<const float4*>(a1 + d));
                                const float4 r2 = __ldg(reinterpret_cast<const float4*>(a2 + d));
                                v[d] = blend(w0, w1, w2, r0.x, r1.x, r2.x);
                                v[d + 1] = blend(w0, w1, w2, r0.y, r1.y, r2.y);
                                v[d + 2] = blend(w0, w1, w2, r0.z, r1.z, r2.z);
                                v[d + 3] = blend(w0, w1, w2, r0.w, r1.w, r2.w);
                            }
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
                            if (d < A) v[d] = blend(w0, w1, w2, __ldg(a0 + d), __ldg(a1 + d), __ldg(a2 + d));
                    }
                    if (ones) vone = blend(w0, w1, w2, 1.0f, 1.0f, 1.0f);
                    if (dep) {                           // view-space z of the corners: shared memory, or the record when the tile took several batches
                        float zv0, zv1, zv2;
                        if (single) { zv0 = s.u.ab.z[0][wli]; zv1 = s.u.ab.z[1][wli]; zv2 = s.u.ab.z[2][wli]; }
                        else {
                            const float* rz = reinterpret_cast<const float*>(recs + fw);
                            zv0 = __ldg(rz + 6); zv1 = __ldg(rz + 7); zv2 = __ldg(rz + 8);
                        }
                        vdep = blend(w0, w1, w2, -zv0, -zv1, -zv2);
                    }
                } else {
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
                improb[gp] = 1.0f;
                imcomp[gp] = 0.0f;
                imidx[gp] = fw + 1;
                s.cnt[ly * TILE + lx] = 255;
            } else {
                imidx[gp] = 0;                   // may be overwritten with the K-th face in phase D
                improb[gp] = 0.0f;               // empty product; phase D overwrites the pixels it reaches
                imcomp[gp] = 1.0f;
                unc = true;
            }
            // every channel goes to its own (tensor, slot): the tables unroll, no group loop at run time
            const int nv = FUSED ? P.va.dim : D;
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) {
                if (d < nv) {
                    P.chan_out[d][px * P.chan_stride[d]] = v[d];
                    if ((P.min_mask >> d) & 1u) vmin = fminf(vmin, v[d]);
                }
            }
            if (FUSED) {                         // the ones and depth channels sit behind the vertex attributes: indexed tables
                int d = nv;
                if (P.va.flags & 1) {
                    P.chan_out[d][px * P.chan_stride[d]] = vone;
                    if ((P.min_mask >> d) & 1u) vmin = fminf(vmin, vone);
                    d++;
                }
                if (P.va.flags & 2) {
                    P.chan_out[d][px * P.chan_stride[d]] = vdep;
                    if ((P.min_mask >> d) & 1u) vmin = fminf(vmin, vdep);
                }
            }
        }
        if (P.min_group >= 0) {
            const unsigned ov = __reduce_min_sync(full_mask, f2ord(vmin));
            if (lane == 0 && ov != f2ord(3.0e38f)) atomicMin(P.out_min, ov);
        }
    }
    {   // uncovered pixels of the tile, one byte per 8 pixels of a row, for the backward's soft part (a warp holds two tile rows)
        const unsigned ub = __ballot_sync(full_mask, unc);
        if ((lane & 7) == 0) {
            const int row = ty0 + (tid >> 4), bytecol = (tx0 >> 3) + ((lane >> 3) & 1);
            if (row < P.height && bytecol * 8 < P.width) {
                const size_t o8 = ((size_t)b * P.height + row) * ((P.width + 7) >> 3) + bytecol;
                P.open8[o8] = (unsigned char)((ub >> lane) & 0xffu);
                P.closed8[o8] = 0;                           // phase D overwrites it where a pixel gets its K-th face
            }
        }
    }
    const int tile_unc = __syncthreads_or(unc ? 1 : 0);        // also: cnt[] complete, the z-buffer is dead
    PHASE_MARK(3);
    if (!tile_unc || P.knum <= 0) return;
#ifdef DIBR_EXP_NOSOFT
    return;
#endif

    // ---- phase D: soft silhouette.  Thread = pixel, block-major: warp w owns the 8x4 block w ---------------------
    const float zscale = (float)P.delta / ((float)P.multiplier * (float)P.multiplier);
    const float sentinel = 4.0f * (float)P.multiplier * (float)P.multiplier;
    const int knum = P.knum;
    const int bx = (warp % NBX) * BW, by = (warp / NBX) * BH;
    const int lx = bx + (lane & 7), ly = by + (lane >> 3);
    const bool valid = (lx < tw) && (ly < th);
    const int pix = ly * TILE + lx;
    const size_t gpix = (size_t)(ty0 + ly) * P.width + (tx0 + lx);
    int c = valid ? (int)s.cnt[pix] : 255;                      // faces accepted so far (255: covered)
    const bool had = (c < knum);
    float q = 0.f, cc = 1.f;                                    // running 1 - prod(1-p) and prod(1-p)
    int wpos = 0;
    for (int batch = 0;; batch++) {
        if (!single) {
            if (wpos >= T.nw) break;
            wpos = fill_list<FUSED>(s, P, f_lo, wpos, T, false, parity);
        }
        const int lcount = s.lcount;
        for (int i = tid; i < LCAP / 4; i += FWD_THREADS) reinterpret_cast<unsigned int*>(s.soft_used)[i] = 0u;
        // per listed face s.smask holds the tile's columns (bits 0-15) and rows (bits 16-31) whose pixel centres lie inside
        // its expanded bbox (rasterizer.py:54-57): written with the records by fill_list
        unsigned int* const smask = s.smask;
        // (no barrier: soft_used is first written behind the collect step's barrier, the lists are stable since fill_list's last one)
        const int c_start = c;
        // passes of HITCAP hits per pixel (one pass unless K > HITCAP)
        for (int skip = 0;; skip += HITCAP) {
            // (1) collect, in ascending face order
            int nh = 0, seen = 0;
            bool more = false;                                  // hits beyond this pass's window
            bool open = (c_start < knum);                       // may still accept in this pass
            const unsigned open32 = __ballot_sync(full_mask, open);
            if (open32) {
                for (int i0 = 0; i0 < lcount; i0 += 32) {
                    const int li = i0 + lane;
                    unsigned m32 = 0u;                          // open pixels of the block inside this lane's face's expanded bbox
                    if (li < lcount) {
                        const unsigned sm = smask[li];
                        const unsigned cm = (sm >> bx) & 0xffu, rm = (sm >> (16 + by)) & 0xfu;
                        m32 = (cm * 0x01010101u) & (((rm * 0x00204081u) & 0x01010101u) * 0xffu) & open32;
                    }
                    unsigned bal = __ballot_sync(full_mask, m32 != 0u);
                    // one accepted (pixel, face) pair of this lane's pixel: list entry i0 + src
                    auto accept = [&](int src) {
                        if (seen >= skip) {
                            if (nh < HITCAP) {
                                s.u.hits[nh][tid] = (unsigned)(i0 + src);
                                nh++;
                                if (c_start + seen + 1 >= knum) {      // the K-th accepted face closes the pixel
                                    open = false;
                                    imidx[gpix] = -(s.lid[i0 + src] + 1);
                                }
                            } else {
                                more = true;
                                open = false;               // nothing more to store in this pass
                            }
                        } else if (c_start + seen + 1 >= knum) {
                            open = false;
                        }
                        if (!more) seen++;
                    };
                    if (__popc(bal) <= 4) {
                        // few faces of this chunk reach the block: hand their masks round one at a time
                        while (bal) {
                            const int src = __ffs(bal) - 1;
                            bal &= bal - 1;
                            const unsigned m = __shfl_sync(full_mask, m32, src);
                            if (open && ((m >> lane) & 1u)) accept(src);
                        }
                    } else {
                        // many: transpose the 32 x 32 bit matrix (5 shuffles), every pixel then walks its own faces only
                        unsigned t = transpose32(m32, lane);
                        if (!open) t = 0u;
                        const int n = __popc(t);
                        // the common chunk: first pass, and no pixel of the block reaches its K-th face or the end of its hit
                        // row inside it -> plain appends, none of accept()'s bookkeeping
                        const bool easy = (skip == 0) && (c_start + seen + n < knum) && (nh + n <= HITCAP);
                        if (__all_sync(full_mask, easy)) {
                            seen += n;
                            for (; t; t &= t - 1) s.u.hits[nh++][tid] = (unsigned)(i0 + __ffs(t) - 1);
                        } else {
                            for (; t && open; t &= t - 1) accept(__ffs(t) - 1);
                        }
                    }
                    if (!__any_sync(full_mask, open)) break;
                }
            }
            // (2) all pairs of the tile, dealt out evenly: every warp publishes the exclusive scan of its pixels' hit counts and
            //     its total BEFORE the barrier that ends the collect step, so no second barrier is needed to find the pairs
            int incl = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(full_mask, incl, o);
                if (lane >= o) incl += t;
            }
            s.E[tid] = (unsigned short)(incl - nh);
            s.nhc[tid] = (unsigned char)nh;
            if (lane == 31) s.warp_tot[parity][warp] = incl;
            const int any_more = __syncthreads_or(more ? 1 : 0);        // hit lists, scans and warp totals visible
            const int par = parity;
            parity ^= 1;
            int total = 0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) total += s.warp_tot[par][w];
            if (total > 0) {
                // ... thread t evaluates the flat range [t*total/256, (t+1)*total/256)
                const int lo = (int)(((long long)tid * total) / FWD_THREADS), hi = (int)(((long long)(tid + 1) * total) / FWD_THREADS);
                if (hi > lo) {
                    // the warp that holds pair `lo` (the last one starting at or before it), then the pixel inside that warp
                    int wq0 = 0, ws = 0;
#pragma unroll
                    for (int w = 0; w < NWARP - 1; w++) {
                        const int t = s.warp_tot[par][w];
                        if (wq0 == w && ws + t <= lo) { ws += t; wq0 = w + 1; }
                    }
                    const int rel = lo - ws;
                    const unsigned short* Ew = s.E + wq0 * 32;
                    int pa = 0;                                 // last lane with Ew[lane] <= rel
#pragma unroll
                    for (int stp = 16; stp >= 1; stp >>= 1) if ((int)Ew[pa + stp] <= rel) pa += stp;
                    int pp = wq0 * 32 + pa, k = rel - (int)Ew[pa];
                    int npp = (int)s.nhc[pp];                   // hits of pixel pp
                    for (int j = lo; j < hi; j++) {
                        while (k >= npp) { pp++; k = 0; npp = (int)s.nhc[pp]; }
                        const int lj = (int)s.u.hits[k][pp];
                        s.soft_used[lj] = 1;                    // benign race: everybody writes 1
                        const float4 a = s.c0[lj];
                        const float2 d = s.c1[lj];
                        const int l = pp & 31, wq = pp >> 5;
                        const SoftHit h = soft_distance(a.x, a.y, a.z, a.w, d.x, d.y,
                                                        s.xs[(wq % NBX) * BW + (l & 7)], s.ys[(wq / NBX) * BH + (l >> 3)], sentinel);
                        s.u.hits[k][pp] = __float_as_uint(soft_prob_enc(h.d2 * zscale));
                        k++;
                    }
                }
                __syncthreads();                                // results (and soft_used) of this pass complete
                // (3) each pixel folds its own results in ascending face order
                for (int k = 0; k < nh; k++) {
                    float p, om;
                    soft_prob_dec(__uint_as_float(s.u.hits[k][tid]), p, om);
                    q = fmaf(p, cc, q);                         // 1 - prod(1-p), accurate for small p
                    cc = cc * om;                               // prod(1-p), accurate for p near 1
                }
            }
            c = min(c_start + seen, 255);
            if (!any_more) break;
            __syncthreads();                                    // the hit lists are rewritten by the next pass
        }
        // soft_used is complete: its writers sit in front of the barrier above the last fold
        // hand the faces that contributed to the backward's work list
        for (int li0 = 0; li0 < lcount; li0 += FWD_THREADS) {
            const int li = li0 + tid;
            if (li < lcount && s.soft_used[li]) mark_face(P, f_lo + s.lid[li], 2u);
        }
        if (single) break;
        // stop early once every uncovered pixel has its K faces
        if (!__syncthreads_or((c < knum) ? 1 : 0)) break;
    }
    if (had) { improb[gpix] = fminf(q, 1.0f); imcomp[gpix] = cc; }     // the recurrence can overshoot 1 by an ulp
    {   // pixels closed by their K-th face (imidx < 0), one byte per row of the block: the backward reads imidx there alone
        const unsigned cb = __ballot_sync(full_mask, had && c >= knum);
        if ((lane & 7) == 0 && cb) {
            const int row = ty0 + ly, bytecol = (tx0 + bx) >> 3;
            if (row < P.height && bytecol * 8 < P.width)
                P.closed8[((size_t)b * P.height + row) * ((P.width + 7) >> 3) + bytecol] = (unsigned char)((cb >> lane) & 0xffu);
        }
    }
    PHASE_MARK(4);
}

int launch_forward(const FwdParams& P, cudaStream_t stream)
{
    const size_t smem = sizeof(FwdSmem);
    {   // opt in to > 48 KB of dynamic shared memory once per device (the call costs ~2 us of host time between launches)
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            e = cudaFuncSetAttribute(dibr_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            e = cudaFuncSetAttribute(dibr_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    const int ntiles = ((P.width + TILE - 1) / TILE) * ((P.height + TILE - 1) / TILE) * P.batch;
    if (P.va.fvid) dibr_forward_kernel<true><<<ntiles, FWD_THREADS, smem, stream>>>(P);
    else dibr_forward_kernel<false><<<ntiles, FWD_THREADS, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
