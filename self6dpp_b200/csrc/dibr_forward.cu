// Forward pass of the B200 DIB-R rasterizer: one CTA (256 threads) per 32x32 screen tile, ~54 KB of
// shared memory so four CTAs share an SM.
//
//   phase A  stream the image's face bboxes through shared memory with TMA bulk copies
//            (cp.async.bulk + mbarrier, double buffered) and keep, in ascending face order, the
//            faces whose EXPANDED bbox touches the tile (warp-ballot compaction)       [binning]
//   phase B  compact the front faces whose bbox really holds a pixel centre of the tile, then
//            face-parallel coverage with 4 lanes per face: barycentric solve in the frozen fp32
//            order and a 64-bit shared-memory atomicMax on (orderable z | ~face id).  The winner
//            is the face with the largest z and, on ties, the smallest index -- what the
//            reference's ascending loop with a strict '>' produces, independent of traversal order.
//   phase C  resolve: per pixel recompute the winner's weights, interpolate the D attributes and
//            write every output tensor / improb=1 / imidx (one image row per warp: coalesced)
//   phase D  soft silhouette for uncovered pixels, one 8x4 pixel block per warp: (1) collect per
//            pixel the first K listed faces whose expanded bbox holds the pixel (ballot pre-filter
//            against the block, per-lane hit lists in shared memory), (2) evaluate
//            exp(-delta d^2 / m^2) for the collected faces with all lanes busy.
//
// Replaces kaolin v0.1's dr_cuda_forward_render_batch + dr_cuda_forward_prob_batch, which the
// reference calls at lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 (one thread
// per pixel looping over ALL faces).  Instead of the reference's five
// B x H x W x 30 scratch tensors (rasterizer.py:144-148) only the K-th accepted face id is kept,
// folded into imidx (see include/dibr_b200.h).
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

#ifdef DIBR_PHASE_TIMING
__device__ unsigned long long g_phase[8];
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(t_ - t_phase)); t_phase = t_; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

constexpr int NWARP = FWD_THREADS / 32;
constexpr int HITCAP = 24;              // collected faces per pixel per round of phase D
constexpr int BW = 8, BH = 4;           // pixel block of one warp in phase D
constexpr int NBX = TILE / BW, NBLK = NBX * (TILE / BH);
static_assert(TILE == 32 && LCAP <= 1024, "rlist packing assumes 5-bit pixel coordinates and 10-bit list indices");
static_assert(NWARP % 4 == 0 && NWARP * 128 * sizeof(float2) <= 2 * SCAN_CHUNK * sizeof(float4), "soft scratch must fit in aux");

struct FwdSmem {
    unsigned long long zkey[TILE * TILE];       //  8 KB
    float4 lbox[LCAP];                          // 16 KB  expanded bbox of listed faces
    int lid[LCAP];                              //  4 KB  local face ids, ascending
    union {                                     // 12 KB
        struct {
            float4 stage[2][SCAN_CHUNK];        //        TMA landing buffers (phase A)
            unsigned int rlist[LCAP];           //        raster candidates (phase B)
        } ab;
        unsigned short hits[HITCAP][FWD_THREADS];   //    per-lane collected list entries (phase D)
    } u;
    float4 aux[2 * SCAN_CHUNK];                 //  8 KB  phase A: TMA stages 2,3 (first pass); phase D: per-warp scratch
    unsigned short sublist[NSUB][SUBCAP];       //  4 KB
    unsigned char cnt[TILE * TILE];             //  1 KB  accepted faces per pixel (255 = covered)
    int subcnt[NSUB];
    int big[BIGCAP];
    float xs[TILE], ys[TILE];
    int warp_tot[2 * NWARP];
    alignas(16) unsigned short warp_cnt16[2][NWARP];   // per-warp hit counts of one scan chunk, read back as two 64-bit words
    int nbig, lcount, next_block, pad0;
    unsigned int sub_uncovered;
    unsigned int unc_blocks;                    // bit (by*4+bx): 8x8 block holds an uncovered pixel
    unsigned char soft_used[LCAP];              // listed faces that entered some pixel's soft product
    uint64_t bar[4];
};

// The first time a face is seen doing `bit`-type work (1: won a pixel, 2: entered a soft product) it is appended
// to the matching work list of the backward.  Warp-aggregated: one counter atomic per warp, every lane must call.
__device__ __forceinline__ void mark_faces_warp(const FwdParams& P, bool want, int g, unsigned bit) {
    bool isnew = false;
    if (want && (__ldcg(&P.face_flags[g]) & bit) == 0u) isnew = (atomicOr(&P.face_flags[g], bit) & bit) == 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, isnew);
    if (bal == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(bal) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&P.list_counts[bit == 1u ? 0 : 1], __popc(bal));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (isnew) (bit == 1u ? P.color_list : P.soft_list)[base + __popc(bal & ((1u << lane) - 1u))] = g;
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

// Phase A: append to the list, in ascending order, the faces in [pos, fnum) whose expanded bbox
// touches the tile, until the list is full.  Returns the next unread face.  Uniform across the CTA.
// The bbox array is streamed through a ring of `nst` TMA stages (256 faces = 4 KB each): a bulk copy takes
// ~1 us to land, so several must be in flight for the scan not to be latency-bound.
__device__ __forceinline__ float4* stage_ptr(FwdSmem& s, int k) {
    return (k < 2) ? &s.u.ab.stage[k][0] : &s.aux[(k - 2) * SCAN_CHUNK];
}

__device__ int fill_list(FwdSmem& s, const float4* __restrict__ bbox, int pos, int fnum, float ex,
                         float tx_lo, float tx_hi, float ty_lo, float ty_hi, uint32_t& phases, int nst)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lcount = 0;
    // prologue: fill the ring
    if (tid == 0) {
        for (int k = 0; k < nst; k++) {
            const int p0 = pos + k * SCAN_CHUNK;
            if (p0 < fnum) {
                const int nf = min(SCAN_CHUNK, fnum - p0);
                mbar_arrive_expect_tx(&s.bar[k], nf * 16);
                tma_load_1d(stage_ptr(s, k), bbox + p0, nf * 16, &s.bar[k]);
            }
        }
    }
    int buf = 0, par = 0;
    bool stopped = false;
    while (pos < fnum) {
        const int nf = min(SCAN_CHUNK, fnum - pos);
        const int npos = pos + nf;
        mbar_wait(&s.bar[buf], (phases >> buf) & 1u);
        phases ^= (1u << buf);

        bool hit = false;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < nf) {
            bb = stage_ptr(s, buf)[tid];
            bb.x -= ex; bb.y -= ex; bb.z += ex; bb.w += ex;        // rasterizer.py:55-57
            // some pixel centre of the tile passes xmin <= x0 < xmax and ymin <= y0 < ymax
            hit = (bb.x <= tx_hi) && (bb.z > tx_lo) && (bb.y <= ty_hi) && (bb.w > ty_lo);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s.warp_cnt16[par][warp] = (unsigned short)__popc(bal);
        __syncthreads();            // counts visible; everybody has read stage[buf]  (the only barrier per chunk)
        // the per-warp counts sit in NWARP/4 64-bit words (16-bit fields): prefix and total by field-wise multiply
        const unsigned long long* wc = reinterpret_cast<const unsigned long long*>(s.warp_cnt16[par]);
        constexpr unsigned long long ONES = 0x0001000100010001ull;
        int base = lcount, tot = 0;
#pragma unroll
        for (int k = 0; k < NWARP / 4; k++) {
            const unsigned long long wv = wc[k];
            const int sk = (int)((wv * ONES) >> 48);
            if (k < (warp >> 2)) base += sk;
            else if (k == (warp >> 2)) base += (int)(((wv & ((1ull << (16 * (warp & 3))) - 1ull)) * ONES) >> 48);
            tot += sk;
        }
        const bool fits = (lcount + tot <= LCAP);
        if (fits && hit) {
            const int slot = base + __popc(bal & ((1u << lane) - 1u));
            s.lbox[slot] = bb;
            s.lid[slot] = pos + tid;
        }
        if (!fits) { stopped = true; break; }
        // refill this stage with the chunk nst ahead (everybody is past reading it: see the barrier above)
        const int p2 = pos + nst * SCAN_CHUNK;
        if (tid == 0 && p2 < fnum) {
            const int nf2 = min(SCAN_CHUNK, fnum - p2);
            mbar_arrive_expect_tx(&s.bar[buf], nf2 * 16);
            tma_load_1d(stage_ptr(s, buf), bbox + p2, nf2 * 16, &s.bar[buf]);
        }
        lcount += tot;
        pos = npos;
        buf = (buf + 1 == nst) ? 0 : buf + 1;
        par ^= 1;                   // the counts are double buffered: a writer of parity p has passed the barrier of
                                    // the previous chunk, which every reader of the older parity-p counts reached after reading
    }
    if (stopped) {
        // the chunk at `pos` stays for the next batch.  Drain every copy still in flight so the barrier phases
        // stay in step: stages buf+1 .. buf+nst-1 hold chunks pos+SCAN_CHUNK .. (the one at `buf` was consumed).
        for (int k = 1; k < nst; k++) {
            const int b2 = (buf + k) % nst;
            if (pos + k * SCAN_CHUNK < fnum) {
                mbar_wait(&s.bar[b2], (phases >> b2) & 1u);
                phases ^= (1u << b2);
            }
        }
    }
    if (tid == 0) s.lcount = lcount;
    __syncthreads();
    return pos;
}

__device__ __forceinline__ void raster_pixel(FwdSmem& s, const FaceK& fk, int f, int lx, int ly) {
    float w0, w1, w2;
    if (!bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2)) return;
    float z0 = blend(w0, w1, w2, fk.az, fk.bz, fk.cz);
    if (!(z0 > -1000.0f)) return;                 // "z0 <= znow" against the initial depth -1000
    z0 = z0 + 0.0f;                               // -0 -> +0 so equal depths compare equal
    const unsigned long long key = ((unsigned long long)f2ord(z0) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)f);
    atomicMax(&s.zkey[ly * TILE + lx], key);
}

struct RasterEntry { int li, c0, nc, r0, nr; };
__device__ __forceinline__ RasterEntry unpack_entry(unsigned int p) {
    RasterEntry e;
    e.li = p & 1023; e.c0 = (p >> 10) & 31; e.nc = ((p >> 15) & 31) + 1; e.r0 = (p >> 20) & 31; e.nr = ((p >> 25) & 31) + 1;
    return e;
}

__device__ __forceinline__ void raster_face_cta(FwdSmem& s, const FaceRec* __restrict__ recs, unsigned int packed) {
    const RasterEntry e = unpack_entry(packed);
    const int f = s.lid[e.li];
    const FaceRec r = recs[f];
    const FaceK fk = make_facek(r);
    for (int i = threadIdx.x; i < e.nc * e.nr; i += FWD_THREADS) raster_pixel(s, fk, f, e.c0 + i % e.nc, e.r0 + i / e.nc);
}

// Phase B
__device__ void raster_list(FwdSmem& s, const FaceRec* __restrict__ recs, int tw, int th, float inv_dx, float inv_dy)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lcount = s.lcount;
    // ---- B1: front faces with a non-empty pixel range -> rlist (packed: list index | c0 | nc-1 | r0 | nr-1).
    //      Two entries per thread per round, their record loads issued together (the records come from L2).
    int rcount = 0;
    for (int i0 = 0; i0 < lcount; i0 += 2 * FWD_THREADS) {
        float4 t2[2], bb[2];
        int idx[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            idx[u] = i0 + u * FWD_THREADS + tid;
            if (idx[u] < lcount) {
                const float4* rp = reinterpret_cast<const float4*>(recs + s.lid[idx[u]]);
                t2[u] = __ldg(rp + 2);                              // cz, nz, -, -
                bb[u] = __ldg(rp + 3);                              // xmin, ymin, xmax, ymax
            }
        }
        unsigned int packed[2] = {0u, 0u};
        bool keep[2] = {false, false};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (idx[u] < lcount && t2[u].y >= 0.0f) {               // front face (K1 culls normalz < 0)
                const int c0 = col_first_ge(s.xs, tw, bb[u].x, inv_dx), c1 = col_first_ge(s.xs, tw, bb[u].z, inv_dx);
                const int r0 = row_first_lt(s.ys, th, bb[u].w, inv_dy), r1 = row_first_lt(s.ys, th, bb[u].y, inv_dy);
                if (c1 > c0 && r1 > r0) {
                    keep[u] = true;
                    packed[u] = (unsigned)idx[u] | ((unsigned)c0 << 10) | ((unsigned)(c1 - c0 - 1) << 15) |
                                ((unsigned)r0 << 20) | ((unsigned)(r1 - r0 - 1) << 25);
                }
            }
        }
        const unsigned bal0 = __ballot_sync(0xffffffffu, keep[0]), bal1 = __ballot_sync(0xffffffffu, keep[1]);
        if (lane == 0) { s.warp_tot[warp] = __popc(bal0); s.warp_tot[NWARP + warp] = __popc(bal1); }
        __syncthreads();
        int base0 = rcount, tot0 = 0, base1 = 0, tot1 = 0;
#pragma unroll
        for (int w = 0; w < NWARP; w++) {
            const int a = s.warp_tot[w], c = s.warp_tot[NWARP + w];
            if (w < warp) { base0 += a; base1 += c; }
            tot0 += a; tot1 += c;
        }
        base1 += rcount + tot0;
        if (keep[0]) s.u.ab.rlist[base0 + __popc(bal0 & ((1u << lane) - 1u))] = packed[0];
        if (keep[1]) s.u.ab.rlist[base1 + __popc(bal1 & ((1u << lane) - 1u))] = packed[1];
        rcount += tot0 + tot1;
        __syncthreads();
    }
    // ---- B2: 4 lanes per face; faces with many pixels in the tile are deferred to the whole CTA
    const int q = tid >> 2, ql = tid & 3;
    for (int e = q; e < rcount; e += FWD_THREADS / 4) {
        const unsigned int packed = s.u.ab.rlist[e];
        const RasterEntry en = unpack_entry(packed);
        const int npx = en.nc * en.nr;
        if (npx > BIG_AREA) {
            if (ql == 0) {
                const int slot = atomicAdd(&s.nbig, 1);
                if (slot < BIGCAP) s.big[slot] = (int)packed;      // beyond BIGCAP: picked up by the rescan below
            }
            continue;
        }
        const int f = s.lid[en.li];
        const FaceRec r = recs[f];
        const FaceK fk = make_facek(r);
        const unsigned inv = 65536u / (unsigned)en.nc + 1u;        // exact i / nc for i < 1024
        for (int i = ql; i < npx; i += 4) {
            const int row = (int)(((unsigned)i * inv) >> 16);
            raster_pixel(s, fk, f, en.c0 + (i - row * en.nc), en.r0 + row);
        }
    }
    __syncthreads();
    // ---- B3: large faces, all threads per face
    const int nbig_all = s.nbig;
    for (int j = 0; j < min(nbig_all, BIGCAP); j++) raster_face_cta(s, recs, (unsigned)s.big[j]);
    if (nbig_all > BIGCAP) {
        for (int e = 0; e < rcount; e++) {
            const unsigned int packed = s.u.ab.rlist[e];
            const RasterEntry en = unpack_entry(packed);
            if (en.nc * en.nr <= BIG_AREA) continue;
            bool listed = false;
            for (int j = 0; j < BIGCAP; j++) listed |= ((unsigned)s.big[j] == packed);
            if (!listed) raster_face_cta(s, recs, packed);
        }
    }
    __syncthreads();
    if (tid == 0) s.nbig = 0;
    __syncthreads();
}

// Phase D for one batch of listed faces
__device__ void soft_list(FwdSmem& s, const FwdParams& P, int f_lo, const FaceRec* __restrict__ recs, int tw, int th, int knum,
                          float zscale, float sentinel, int* __restrict__ imidx_img, float* __restrict__ improb_img,
                          float* __restrict__ imcomp_img, int width, int tx0, int ty0, bool first_batch)
{
    const unsigned full_mask = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lcount = s.lcount;
    for (int i = tid; i < LCAP / 4; i += FWD_THREADS) reinterpret_cast<unsigned int*>(s.soft_used)[i] = 0u;
    // ---- per-sub-tile (16x16) ordered lists: warp w builds sub-tile w -----------------------------------
    if (warp < NSUB) {
        const int st = warp;
        const int sx = (st % (TILE / SUB)) * SUB, sy = (st / (TILE / SUB)) * SUB;
        int n = 0;
        if (((s.sub_uncovered >> st) & 1u) && sx < tw && sy < th) {
            const float x_lo = s.xs[sx], x_hi = s.xs[min(sx + SUB, tw) - 1];
            const float y_hi = s.ys[sy], y_lo = s.ys[min(sy + SUB, th) - 1];
            for (int i0 = 0; i0 < lcount; i0 += 32) {
                const int i = i0 + lane;
                bool hit = false;
                if (i < lcount) {
                    const float4 bb = s.lbox[i];
                    hit = (bb.x <= x_hi) && (bb.z > x_lo) && (bb.y <= y_hi) && (bb.w > y_lo);
                }
                const unsigned bal = __ballot_sync(full_mask, hit);
                if (hit) {
                    const int slot = n + __popc(bal & ((1u << lane) - 1u));
                    if (slot < SUBCAP) s.sublist[st][slot] = (unsigned short)i;
                }
                n += __popc(bal);
            }
        }
        if (lane == 0) s.subcnt[st] = n;          // n > SUBCAP: overflow, fall back to the full list
    }
    __syncthreads();
    // ---- 8x4 pixel blocks, handed out dynamically ---------------------------------------------------------
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(&s.next_block, 1);
        blk = __shfl_sync(full_mask, blk, 0);
        if (blk >= NBLK) break;
        const int bx = (blk % NBX) * BW, by = (blk / NBX) * BH;
        if (bx >= tw || by >= th) continue;
        const int lx = bx + (lane & 7), ly = by + (lane >> 3);
        const bool valid = (lx < tw) && (ly < th);
        const int pix = ly * TILE + lx;
        int c = valid ? (int)s.cnt[pix] : 255;
        bool open = (c < knum);                    // uncovered and not yet saturated
        if (!__any_sync(full_mask, open)) continue;
        const int st = (by / SUB) * (TILE / SUB) + (bx / SUB);
        const int sn = s.subcnt[st];
        if (sn == 0) continue;
        const bool whole = (sn > SUBCAP);
        const int n = whole ? lcount : sn;
        const float x0 = valid ? s.xs[lx] : 0.f, y0 = valid ? s.ys[ly] : 0.f;
        // pre-filter rectangle = bounding box of the block's OPEN pixels (they only ever close, so it stays valid)
        const unsigned colmask = __reduce_or_sync(full_mask, open ? (1u << (lane & 7)) : 0u);
        const unsigned rowmask = __reduce_or_sync(full_mask, open ? (1u << (lane >> 3)) : 0u);
        const float x_lo = s.xs[bx + __ffs(colmask) - 1], x_hi = s.xs[bx + 31 - __clz(colmask)];
        const float y_hi = s.ys[by + __ffs(rowmask) - 1], y_lo = s.ys[by + 31 - __clz(rowmask)];
        const bool had = open;
        const size_t gpix = (size_t)(ty0 + ly) * width + (tx0 + lx);
        // running 1 - prod(1-p) and prod(1-p): start of the product, or what earlier list batches left in the images
        float q = (had && !first_batch) ? improb_img[gpix] : 0.f, cc = (had && !first_batch) ? imcomp_img[gpix] : 1.f;
        int nh = 0;                                // collected, not yet evaluated

        // (2) evaluate the collected (pixel, face) pairs.  Lanes hold very different numbers of hits, so the pairs
        // of four k-levels at a time are flattened (warp scan) and dealt out evenly: every lane evaluates one
        // pair per pass; then each pixel folds ITS results in ascending face order.  Scratch lives in the z-buffer
        // (dead after phase C) and in aux.
        float2* const res = reinterpret_cast<float2*>(s.aux) + warp * 128;                      // 1 KB of aux
        unsigned short* const prs = reinterpret_cast<unsigned short*>(s.zkey) + warp * 128;     // 256 B of the dead z-buffer
        auto flush_hits = [&]() {
            __syncwarp();                          // the hit lists were written by other lanes
            const int kmax = __reduce_max_sync(full_mask, nh);
            for (int k0 = 0; k0 < kmax; k0 += 4) {
                const int cnt = min(max(nh - k0, 0), 4);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(full_mask, incl, o);
                    if (lane >= o) incl += t;
                }
                const int excl = incl - cnt;
                const int T = __shfl_sync(full_mask, incl, 31);
                for (int i = 0; i < cnt; i++) prs[excl + i] = (unsigned short)((lane << 8) | (k0 + i));
                __syncwarp();
                for (int j = lane; j < T; j += 32) {
                    const int pk = prs[j];
                    const int l = pk >> 8, k = pk & 255;
                    const int lj = s.u.hits[k][warp * 32 + l];
                    s.soft_used[lj] = 1;           // benign race: everybody writes 1
                    const float4* rp = reinterpret_cast<const float4*>(recs + s.lid[lj]);
                    const float4 g0 = __ldg(rp), g1 = __ldg(rp + 1);
                    const SoftHit h = soft_distance(g0.x, g0.y, g0.z, g0.w, g1.x, g1.y,
                                                    s.xs[bx + (l & 7)], s.ys[by + (l >> 3)], sentinel);
                    float p, om;
                    soft_prob(h.d2 * zscale, p, om);
                    res[j] = make_float2(p, om);
                }
                __syncwarp();
                for (int i = 0; i < cnt; i++) {
                    const float2 r = res[excl + i];
                    q = fmaf(r.x, cc, q);          // 1 - prod(1-p), accurate for small p
                    cc = cc * r.y;                 // prod(1-p), accurate for p near 1
                }
                __syncwarp();
            }
            nh = 0;
        };

        // (1) collect, in ascending face order
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            int li = -1;
            bool hit = false;
            if (i < n) {
                li = whole ? i : (int)s.sublist[st][i];
                const float4 bb = s.lbox[li];
                hit = (bb.x <= x_hi) && (bb.z > x_lo) && (bb.y <= y_hi) && (bb.w > y_lo);
            }
            unsigned bal = __ballot_sync(full_mask, hit);
            const int nb = __popc(bal);
            if (nb == 0) continue;
            if (__reduce_max_sync(full_mask, nh) + nb > HITCAP) flush_hits();
            const bool careful = nb > HITCAP;      // more candidates than a lane can hold even when empty
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const int lj = __shfl_sync(full_mask, li, src);
                const float4 bb = s.lbox[lj];
                if (open && !(x0 < bb.x || x0 >= bb.z || y0 < bb.y || y0 >= bb.w)) {
                    s.u.hits[nh][tid] = (unsigned short)lj;
                    nh++;
                    c++;
                    if (c >= knum) {               // the K-th accepted face closes the pixel
                        open = false;
                        imidx_img[gpix] = -(s.lid[lj] + 1);
                    }
                }
                if (careful && __any_sync(full_mask, nh >= HITCAP)) flush_hits();
            }
            if (!__any_sync(full_mask, open)) break;
        }
        flush_hits();
        if (had) { improb_img[gpix] = fminf(q, 1.0f); imcomp_img[gpix] = cc; s.cnt[pix] = (unsigned char)c; }   // the recurrence can overshoot 1 by an ulp
    }
    __syncthreads();
    if (tid == 0) s.next_block = 0;
    // hand the faces that contributed to the backward's work list
    for (int li0 = 0; li0 < lcount; li0 += FWD_THREADS) {
        const int li = li0 + tid;
        const bool used = (li < lcount) && s.soft_used[li];
        mark_faces_warp(P, used, used ? f_lo + s.lid[li] : 0, 2u);
    }
    __syncthreads();
}

// zero-fill a full 32x32 tile of one [H,W,CH] image with 128-bit stores: CH stores per thread, no divisions at run time
template <int CH>
__device__ __forceinline__ void zero_full_tile(float* __restrict__ img, int width, int tx0, int ty0)
{
    constexpr int RV = TILE * CH / 4;                    // float4 per tile row
#pragma unroll
    for (int i0 = 0; i0 < TILE * RV; i0 += FWD_THREADS) {
        const int i = i0 + threadIdx.x;
        if (i < TILE * RV) {
            const int r = i / RV, c = i - r * RV;
            reinterpret_cast<float4*>(img + ((size_t)(ty0 + r) * width + tx0) * CH)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// zero-fill rows [0,th) x [0,tw) of one [H,W,ch] image tile
__device__ __forceinline__ void zero_tile(float* __restrict__ img, int width, int ch, int tx0, int ty0, int tw, int th)
{
    const bool aligned = (((size_t)width * ch) & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (tw == TILE && th == TILE && aligned && ch <= 4) {          // tx0 is a multiple of 32: rows start 16 B aligned
        switch (ch) {
            case 1: zero_full_tile<1>(img, width, tx0, ty0); return;
            case 2: zero_full_tile<2>(img, width, tx0, ty0); return;
            case 3: zero_full_tile<3>(img, width, tx0, ty0); return;
            default: zero_full_tile<4>(img, width, tx0, ty0); return;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rowf = tw * ch;
    for (int r = warp; r < th; r += NWARP) {
        float* row = img + ((size_t)(ty0 + r) * width + tx0) * ch;
        for (int c = lane; c < rowf; c += 32) row[c] = 0.f;
    }
}

#ifndef DIBR_FWD_MIN_CTAS
#define DIBR_FWD_MIN_CTAS (1024 / DIBR_FWD_THREADS)
#endif
__global__ void __launch_bounds__(FWD_THREADS, DIBR_FWD_MIN_CTAS)
dibr_forward_kernel(FwdParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    const int tw = min(TILE, P.width - tx0), th = min(TILE, P.height - ty0);
    const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
    const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
    const int fnum = f_hi - f_lo;
    const FaceRec* __restrict__ recs = P.recs + f_lo;
    const float4* __restrict__ bbox = P.bbox + f_lo;
    const int D = P.num_attr;
    const size_t img_pix = (size_t)b * P.height * P.width;
    float* __restrict__ improb = P.improb + img_pix;
    float* __restrict__ imcomp = P.imcomp + img_pix;
    int* __restrict__ imidx = P.imidx + img_pix;
    const float ex = P.expand_mul;

    // ---- whole-image cull (imgbox: ordered maxima of (-xmin,-ymin,xmax,ymax) over the image's faces) ----------
    const float t_xlo = pix_x(tx0, P.width, P.multiplier), t_xhi = pix_x(tx0 + tw - 1, P.width, P.multiplier);
    const float t_yhi = pix_y(ty0, P.height, P.multiplier), t_ylo = pix_y(ty0 + th - 1, P.height, P.multiplier);
    bool touched = false;
    if (fnum > 0) {
        const uint4 ib = P.imgbox[b];
        const float ixmin = -ord2f(ib.x), iymin = -ord2f(ib.y), ixmax = ord2f(ib.z), iymax = ord2f(ib.w);
        touched = (ib.z != 0u) && (ixmin - ex <= t_xhi) && (ixmax + ex > t_xlo) && (iymin - ex <= t_yhi) && (iymax + ex > t_ylo);
    }
    unsigned short* __restrict__ unc_out = P.unc_blocks + ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (!touched) {          // nothing near this tile: zeros everywhere (imcomp = 1: empty product)
        if (tid == 0) *unc_out = 0xffffu;
        if (tid == 0 && P.min_group >= 0) atomicMin(P.out_min, f2ord(0.0f));
        for (int g = 0; g < P.n_out; g++) zero_tile(P.out[g] + img_pix * P.out_ch[g], P.width, P.out_ch[g], tx0, ty0, tw, th);
        zero_tile(improb, P.width, 1, tx0, ty0, tw, th);
        zero_tile(reinterpret_cast<float*>(imidx), P.width, 1, tx0, ty0, tw, th);
        for (int r = tid >> 5; r < th; r += NWARP)
            for (int c = tid & 31; c < tw; c += 32) imcomp[(size_t)(ty0 + r) * P.width + tx0 + c] = 1.0f;
        return;
    }

#ifdef DIBR_PHASE_TIMING
    long long t_phase = clock64();
    if (tid == 0) atomicAdd(&g_phase[7], 1ull);
#endif
    // ---- tile set-up ----------------------------------------------------------------------------
    if (tid < TILE) {
        s.xs[tid] = (tid < tw) ? pix_x(tx0 + tid, P.width, P.multiplier) : 3.0e38f;
    } else if (tid < 2 * TILE) {
        const int r = tid - TILE;
        s.ys[r] = (r < th) ? pix_y(ty0 + r, P.height, P.multiplier) : -3.0e38f;
    }
    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        mbar_init(&s.bar[2], 1);
        mbar_init(&s.bar[3], 1);
        mbar_fence_init();
        s.nbig = 0; s.next_block = 0; s.lcount = 0; s.sub_uncovered = 0u; s.unc_blocks = 0u;
    }
    for (int i = tid; i < TILE * TILE; i += FWD_THREADS) {
        s.zkey[i] = 0ull; s.cnt[i] = 0;
    }
    __syncthreads();
    const float tx_lo = s.xs[0], tx_hi = s.xs[tw - 1];
    const float ty_hi = s.ys[0], ty_lo = s.ys[th - 1];
    const float inv_dx = 0.5f * (float)P.width / (float)P.multiplier;     // pixel pitch is 2m/W
    const float inv_dy = 0.5f * (float)P.height / (float)P.multiplier;
    uint32_t phases = 0;          // one parity bit per TMA stage barrier

    int nbatch = 0;
    {
        int pos = 0;
        PHASE_MARK(0);
        while (pos < fnum) {
            pos = fill_list(s, bbox, pos, fnum, ex, tx_lo, tx_hi, ty_lo, ty_hi, phases, 4);
            PHASE_MARK(1);
            if (s.lcount > 0) raster_list(s, recs, tw, th, inv_dx, inv_dy);
            PHASE_MARK(2);
            nbatch++;
        }
    }

    // ---- phase C: resolve (one image row per warp) ---------------------------------------------------------
    const float* __restrict__ fattr = P.face_attr + (size_t)f_lo * 3 * D;
    bool any_unc = false;
    float vmin = 3.0e38f;                        // running minimum of output group P.min_group
    // two image rows per pass: the winners' 2-D corners (32 B of the record) for both pixels are requested before
    // either is used, then the face-flag words of both winners -- every one of these comes from L2
#pragma unroll 1
    for (int it = 0; it < (TILE * TILE) / FWD_THREADS; it += 2) {
        const int lx = tid & 31;
        int lys[2], fw[2];
        float4 c0[2], c1[2];
        bool val[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            lys[u] = (it + u) * NWARP + (tid >> 5);
            val[u] = (lx < tw) && (lys[u] < th);
            fw[u] = -1;
            if (val[u]) {
                const unsigned long long key = s.zkey[lys[u] * TILE + lx];
                if (key != 0ull) {
                    fw[u] = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
                    const float4* rp = reinterpret_cast<const float4*>(recs + fw[u]);
                    c0[u] = __ldg(rp);
                    c1[u] = __ldg(rp + 1);
                }
            }
        }
        unsigned fl[2];
        bool lead[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            // winners go on the backward's colour work list (run-length de-duplicated along the row)
            const int prev = __shfl_up_sync(0xffffffffu, fw[u], 1);
            lead[u] = fw[u] >= 0 && (lx == 0 || prev != fw[u]);
            fl[u] = lead[u] ? __ldcg(&P.face_flags[f_lo + fw[u]]) : 1u;
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int ly = lys[u];
            bool unc = false;
            if (val[u]) {
                const size_t gp = (size_t)(ty0 + ly) * P.width + (tx0 + lx);
                const size_t px = img_pix + gp;
                if (fw[u] >= 0) {
                    const int f = fw[u];
                    FaceRec r;
                    r.ax = c0[u].x; r.ay = c0[u].y; r.bx = c0[u].z; r.by = c0[u].w; r.cx = c1[u].x; r.cy = c1[u].y;
                    r.az = r.bz = r.cz = 0.f;
                    const FaceK fk = make_facek(r);
                    float w0, w1, w2;
                    bary(fk, s.xs[lx], s.ys[ly], w0, w1, w2);
                    const float* a = fattr + (size_t)f * 3 * D;
                    int base = 0;
                    for (int g = 0; g < P.n_out; g++) {
                        const int ch = P.out_ch[g];
                        float* o = P.out[g] + px * ch;
                        if (ch == 4 && ((D | base) & 3) == 0) {
                            const float4 r0 = __ldg(reinterpret_cast<const float4*>(a + base));
                            const float4 r1 = __ldg(reinterpret_cast<const float4*>(a + D + base));
                            const float4 r2 = __ldg(reinterpret_cast<const float4*>(a + 2 * D + base));
                            float4 v;
                            v.x = blend(w0, w1, w2, r0.x, r1.x, r2.x);
                            v.y = blend(w0, w1, w2, r0.y, r1.y, r2.y);
                            v.z = blend(w0, w1, w2, r0.z, r1.z, r2.z);
                            v.w = blend(w0, w1, w2, r0.w, r1.w, r2.w);
                            *reinterpret_cast<float4*>(o) = v;
                            if (g == P.min_group) vmin = fminf(vmin, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
                        } else {
                            for (int c = 0; c < ch; c++) {
                                const float v = blend(w0, w1, w2, __ldg(a + base + c), __ldg(a + D + base + c), __ldg(a + 2 * D + base + c));
                                o[c] = v;
                                if (g == P.min_group) vmin = fminf(vmin, v);
                            }
                        }
                        base += ch;
                    }
                    improb[gp] = 1.0f;
                    imcomp[gp] = 0.0f;
                    imidx[gp] = f + 1;
                    s.cnt[ly * TILE + lx] = 255;
                } else {
                    for (int g = 0; g < P.n_out; g++) {
                        const int ch = P.out_ch[g];
                        float* o = P.out[g] + px * ch;
                        if (ch == 4) *reinterpret_cast<float4*>(o) = make_float4(0.f, 0.f, 0.f, 0.f);
                        else for (int c = 0; c < ch; c++) o[c] = 0.f;
                    }
                    imidx[gp] = 0;                   // may be overwritten with the K-th face in phase D
                    improb[gp] = 0.0f;               // empty product; phase D overwrites the pixels it reaches
                    imcomp[gp] = 1.0f;
                    unc = true;
                    vmin = fminf(vmin, 0.0f);
                }
            }
            // append first-time winners to the colour list: one counter atomic per warp
            {
                bool isnew = false;
                const int g = f_lo + max(fw[u], 0);
                if (lead[u] && (fl[u] & 1u) == 0u) isnew = (atomicOr(&P.face_flags[g], 1u) & 1u) == 0u;
                const unsigned nb = __ballot_sync(0xffffffffu, isnew);
                if (nb) {
                    const int leader = __ffs(nb) - 1;
                    int lb = 0;
                    if (lx == leader) lb = atomicAdd(&P.list_counts[0], __popc(nb));
                    lb = __shfl_sync(0xffffffffu, lb, leader);
                    if (isnew) P.color_list[lb + __popc(nb & ((1u << lx) - 1u))] = g;
                }
            }
            // which 16x16 sub-tiles / 8x8 blocks still hold uncovered pixels (a warp is one row: lanes 0-15 | 16-31)
            const unsigned bal = __ballot_sync(0xffffffffu, unc);
            if (bal) {
                any_unc = true;
                if (lx == 0) {
                    const int st0 = (ly / SUB) * (TILE / SUB);
                    unsigned m = 0;
                    if (bal & 0x0000ffffu) m |= 1u << st0;
                    if (bal & 0xffff0000u) m |= 1u << (st0 + 1);
                    atomicOr(&s.sub_uncovered, m);
                    unsigned m8 = 0;
                    const int brow = (ly >> 3) * 4;
                    if (bal & 0x000000ffu) m8 |= 1u << brow;
                    if (bal & 0x0000ff00u) m8 |= 1u << (brow + 1);
                    if (bal & 0x00ff0000u) m8 |= 1u << (brow + 2);
                    if (bal & 0xff000000u) m8 |= 1u << (brow + 3);
                    atomicOr(&s.unc_blocks, m8);
                }
            }
        }
    }
    if (P.min_group >= 0) {
        const unsigned ov = __reduce_min_sync(0xffffffffu, f2ord(vmin));
        if ((tid & 31) == 0 && ov != f2ord(3.0e38f)) atomicMin(P.out_min, ov);
    }
    const int tile_unc = __syncthreads_or(any_unc ? 1 : 0);
    PHASE_MARK(3);
    if (tid == 0) *unc_out = (unsigned short)s.unc_blocks;

    // ---- phase D: soft silhouette ------------------------------------------------------------------
    if (tile_unc && P.knum > 0) {
        const float zscale = (float)P.delta / ((float)P.multiplier * (float)P.multiplier);
        const float sentinel = 4.0f * (float)P.multiplier * (float)P.multiplier;
        if (nbatch == 1) {
            if (s.lcount > 0) soft_list(s, P, f_lo, recs, tw, th, P.knum, zscale, sentinel, imidx, improb, imcomp, P.width, tx0, ty0, true);
        } else {
            int pos = 0;
            bool first = true;
            while (pos < fnum) {
                pos = fill_list(s, bbox, pos, fnum, ex, tx_lo, tx_hi, ty_lo, ty_hi, phases, 2);      // aux holds soft scratch
                if (s.lcount > 0) {
                    soft_list(s, P, f_lo, recs, tw, th, P.knum, zscale, sentinel, imidx, improb, imcomp, P.width, tx0, ty0, first);
                    first = false;
                }
                // stop early once every uncovered pixel has its K faces
                bool open = false;
                for (int i = tid; i < TILE * TILE; i += FWD_THREADS) open |= ((int)s.cnt[i] < P.knum);
                if (!__syncthreads_or(open ? 1 : 0)) break;
            }
        }
    }
    PHASE_MARK(4);
}

// out = (n - min) / (||n - min|| + 1e-5) * mask  (renderer_dibr.py:284-285)
__global__ void __launch_bounds__(256) normal_map_kernel(const float* __restrict__ n, const float* __restrict__ mask,
                                                         const unsigned int* __restrict__ min_ordered, float* __restrict__ out, long long npix)
{
    const float mn = ord2f(*min_ordered);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const float a = n[3 * i] - mn, b = n[3 * i + 1] - mn, c = n[3 * i + 2] - mn;
        const float len = sqrtf(a * a + b * b + c * c) + 1e-5f;
        const float m = mask[i];
        out[3 * i] = a / len * m; out[3 * i + 1] = b / len * m; out[3 * i + 2] = c / len * m;
    }
}

int launch_normal_map(const float* n, const float* mask, const unsigned int* min_ordered, float* out, long long npix, cudaStream_t stream)
{
    if (npix == 0) return 0;
    const int grid = (int)((npix + 255) / 256 < 148 * 16 ? (npix + 255) / 256 : 148 * 16);
    normal_map_kernel<<<grid, 256, 0, stream>>>(n, mask, min_ordered, out, npix);
    return (int)cudaGetLastError();
}

#ifdef DIBR_PHASE_TIMING
extern "C" void dibr_debug_phase_cycles(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_phase, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof(z)); }
}
#endif

int launch_forward(const FwdParams& P, cudaStream_t stream)
{
    static bool attr_set = false;
    const size_t smem = sizeof(FwdSmem);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dibr_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((P.width + TILE - 1) / TILE, (P.height + TILE - 1) / TILE, P.batch);
    dibr_forward_kernel<<<grid, FWD_THREADS, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
