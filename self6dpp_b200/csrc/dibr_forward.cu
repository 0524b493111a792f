// Forward pass of the B200 DIB-R rasterizer: one CTA per 64x64 screen tile.
//
//   phase A  stream the image's face bboxes through shared memory with TMA bulk copies
//            (cp.async.bulk + mbarrier, double buffered) and keep, in ascending face order, the
//            faces whose EXPANDED bbox touches the tile (warp-ballot compaction)       [binning]
//   phase B  face-parallel coverage: every listed front face walks the pixel centres inside
//            its bbox, solves the barycentric system in the frozen fp32 order and does a 64-bit
//            shared-memory atomicMax on (orderable z | ~face id): the winner is the face with
//            the largest z and, on ties, the smallest index -- exactly what the reference's
//            ascending loop with a strict '>' produces, independent of traversal order.
//   phase C  resolve: per pixel recompute the winner's weights, interpolate D attributes,
//            write im / improb=1 / imidx (coalesced, 128-bit stores when D % 4 == 0)
//   phase D  soft silhouette for uncovered pixels: per 16x16 sub-tile ordered lists, then per
//            8x4 pixel block (one warp) the first K faces in index order whose expanded bbox
//            holds the pixel contribute exp(-delta d^2 / m^2).
//
// Replaces kaolin v0.1's dr_cuda_forward_render_batch + dr_cuda_forward_prob_batch, which the
// reference calls at lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 (one thread
// per pixel looping over ALL faces).  Instead of the reference's five
// B x H x W x 30 scratch tensors (rasterizer.py:144-148) only the K-th accepted face id is kept,
// folded into imidx (see include/dibr_b200.h).
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

struct FwdSmem {
    unsigned long long zkey[TILE * TILE];   // 32 KB
    float4 lbox[LCAP];                      // 32 KB  expanded bbox of listed faces
    int lid[LCAP];                          //  8 KB  local face ids, ascending
    float4 stage[2][SCAN_CHUNK];            //  8 KB  TMA landing buffers
    float soft_q[TILE * TILE];              // 16 KB  1 - prod(1-p)
    float soft_c[TILE * TILE];              // 16 KB  prod(1-p)
    unsigned short sublist[NSUB][SUBCAP];   // 12 KB
    unsigned char cnt[TILE * TILE];         //  4 KB  accepted faces per pixel (255 = covered)
    int subcnt[NSUB];
    int big[BIGCAP];
    float xs[TILE], ys[TILE];
    int warp_tot[FWD_THREADS / 32];
    int nbig, lcount, next_block, flag;
    unsigned int sub_uncovered;
    uint64_t bar[2];
};

// first index i in [0,n) with v[i] >= x, v ascending (n if none)
__device__ __forceinline__ int lower_asc(const float* v, int n, float x) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] >= x) hi = mid; else lo = mid + 1; }
    return lo;
}
// first index i in [0,n) with v[i] < x, v descending (n if none)
__device__ __forceinline__ int lower_desc(const float* v, int n, float x) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < x) hi = mid; else lo = mid + 1; }
    return lo;
}

// Phase A: append to the list, in ascending order, the faces in [pos, fnum) whose expanded bbox
// touches the tile, until the list is full.  Returns the next unread face.  Uniform across the CTA.
__device__ int fill_list(FwdSmem& s, const float4* __restrict__ bbox, int pos, int fnum, float ex,
                         float tx_lo, float tx_hi, float ty_lo, float ty_hi, uint32_t& phase0, uint32_t& phase1)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lcount = 0;
    // prologue: stage the first chunk
    if (tid == 0 && pos < fnum) {
        const int nf = min(SCAN_CHUNK, fnum - pos);
        mbar_arrive_expect_tx(&s.bar[0], nf * 16);
        tma_load_1d(&s.stage[0][0], bbox + pos, nf * 16, &s.bar[0]);
    }
    int buf = 0;
    while (pos < fnum) {
        const int nf = min(SCAN_CHUNK, fnum - pos);
        const int npos = pos + nf;
        // prefetch the following chunk into the other buffer
        if (tid == 0 && npos < fnum) {
            const int nf2 = min(SCAN_CHUNK, fnum - npos);
            mbar_arrive_expect_tx(&s.bar[buf ^ 1], nf2 * 16);
            tma_load_1d(&s.stage[buf ^ 1][0], bbox + npos, nf2 * 16, &s.bar[buf ^ 1]);
        }
        uint32_t& ph = buf ? phase1 : phase0;
        mbar_wait(&s.bar[buf], ph);
        ph ^= 1;

        bool hit = false;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < nf) {
            bb = s.stage[buf][tid];
            bb.x -= ex; bb.y -= ex; bb.z += ex; bb.w += ex;        // rasterizer.py:55-57
            // some pixel centre of the tile passes xmin <= x0 < xmax and ymin <= y0 < ymax
            hit = (bb.x <= tx_hi) && (bb.z > tx_lo) && (bb.y <= ty_hi) && (bb.w > ty_lo);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s.warp_tot[warp] = __popc(bal);
        __syncthreads();
        int base = lcount, tot = 0;
#pragma unroll
        for (int w = 0; w < FWD_THREADS / 32; w++) {
            const int c = s.warp_tot[w];
            if (w < warp) base += c;
            tot += c;
        }
        const bool fits = (lcount + tot <= LCAP);
        if (fits && hit) {
            const int slot = base + __popc(bal & ((1u << lane) - 1u));
            s.lbox[slot] = bb;
            s.lid[slot] = pos + tid;
        }
        __syncthreads();            // stage[buf] and warp_tot are free again
        if (!fits) {
            // leave this chunk for the next batch; drain the prefetch so the phases stay in step
            if (npos < fnum) {
                uint32_t& ph2 = (buf ^ 1) ? phase1 : phase0;
                mbar_wait(&s.bar[buf ^ 1], ph2);
                ph2 ^= 1;
            }
            __syncthreads();
            break;
        }
        lcount += tot;
        pos = npos;
        buf ^= 1;
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

// Phase B
__device__ void raster_list(FwdSmem& s, const FaceRec* __restrict__ recs, int tw, int th)
{
    const int tid = threadIdx.x;
    const int lcount = s.lcount;
    for (int i = tid; i < lcount; i += FWD_THREADS) {
        const int f = s.lid[i];
        const FaceRec r = recs[f];
        if (r.nz < 0.0f) continue;                 // back face (K1 only)
        const int c0 = lower_asc(s.xs, tw, r.xmin), c1 = lower_asc(s.xs, tw, r.xmax);
        const int r0 = lower_desc(s.ys, th, r.ymax), r1 = lower_desc(s.ys, th, r.ymin);
        const int nc = c1 - c0, nr = r1 - r0;
        if (nc <= 0 || nr <= 0) continue;
        if (nc * nr > BIG_AREA) {
            const int slot = atomicAdd(&s.nbig, 1);
            if (slot < BIGCAP) { s.big[slot] = f; continue; }
        }
        const FaceK fk = make_facek(r);
        for (int ly = r0; ly < r1; ly++)
            for (int lx = c0; lx < c1; lx++) raster_pixel(s, fk, f, lx, ly);
    }
    __syncthreads();
    const int nbig = min(s.nbig, BIGCAP);
    for (int j = 0; j < nbig; j++) {
        const int f = s.big[j];
        const FaceRec r = recs[f];
        const int c0 = lower_asc(s.xs, tw, r.xmin), c1 = lower_asc(s.xs, tw, r.xmax);
        const int r0 = lower_desc(s.ys, th, r.ymax), r1 = lower_desc(s.ys, th, r.ymin);
        const int nc = c1 - c0, npx = nc * (r1 - r0);
        const FaceK fk = make_facek(r);
        for (int i = tid; i < npx; i += FWD_THREADS) raster_pixel(s, fk, f, c0 + i % nc, r0 + i / nc);
    }
    __syncthreads();
    if (tid == 0) s.nbig = 0;
    __syncthreads();
}

// Phase D for one batch of listed faces
__device__ void soft_list(FwdSmem& s, const FaceRec* __restrict__ recs, int tw, int th, int knum,
                          float zscale, float sentinel, int* __restrict__ imidx_img, int width, int tx0, int ty0)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lcount = s.lcount;
    // ---- per-sub-tile ordered lists -----------------------------------------------------------
    for (int st = warp; st < NSUB; st += FWD_THREADS / 32) {
        if (!((s.sub_uncovered >> st) & 1u)) { if (lane == 0) s.subcnt[st] = 0; continue; }
        const int sx = (st % (TILE / SUB)) * SUB, sy = (st / (TILE / SUB)) * SUB;
        if (sx >= tw || sy >= th) { if (lane == 0) s.subcnt[st] = 0; continue; }
        const float x_lo = s.xs[sx], x_hi = s.xs[min(sx + SUB, tw) - 1];
        const float y_hi = s.ys[sy], y_lo = s.ys[min(sy + SUB, th) - 1];
        int n = 0;
        for (int i0 = 0; i0 < lcount; i0 += 32) {
            const int i = i0 + lane;
            bool hit = false;
            if (i < lcount) {
                const float4 bb = s.lbox[i];
                hit = (bb.x <= x_hi) && (bb.z > x_lo) && (bb.y <= y_hi) && (bb.w > y_lo);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int slot = n + __popc(bal & ((1u << lane) - 1u));
                if (slot < SUBCAP) s.sublist[st][slot] = (unsigned short)i;
            }
            n += __popc(bal);
        }
        if (lane == 0) s.subcnt[st] = n;          // n > SUBCAP: overflow, fall back to the full list
    }
    __syncthreads();
    // ---- 8x4 pixel blocks, handed out dynamically ---------------------------------------------
    constexpr int BW = 8, BH = 4, NBX = TILE / BW, NBLK = NBX * (TILE / BH);
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(&s.next_block, 1);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if (blk >= NBLK) break;
        const int bx = (blk % NBX) * BW, by = (blk / NBX) * BH;
        if (bx >= tw || by >= th) continue;
        const int lx = bx + (lane & 7), ly = by + (lane >> 3);
        const bool valid = (lx < tw) && (ly < th);
        const int pix = ly * TILE + lx;
        int c = valid ? (int)s.cnt[pix] : 255;
        bool open = (c < knum);                    // uncovered and not yet saturated
        if (!__any_sync(0xffffffffu, open)) continue;
        const int st = (by / SUB) * (TILE / SUB) + (bx / SUB);
        const int sn = s.subcnt[st];
        if (sn == 0) continue;
        const bool full = (sn > SUBCAP);
        const int n = full ? lcount : sn;
        const float x0 = valid ? s.xs[lx] : 0.f, y0 = valid ? s.ys[ly] : 0.f;
        const float x_lo = s.xs[bx], x_hi = s.xs[min(bx + BW, tw) - 1];
        const float y_hi = s.ys[by], y_lo = s.ys[min(by + BH, th) - 1];
        float q = valid ? s.soft_q[pix] : 0.f, cc = valid ? s.soft_c[pix] : 1.f;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            int li = -1;
            bool hit = false;
            if (i < n) {
                li = full ? i : (int)s.sublist[st][i];
                const float4 bb = s.lbox[li];
                hit = (bb.x <= x_hi) && (bb.z > x_lo) && (bb.y <= y_hi) && (bb.w > y_lo);
            }
            unsigned bal = __ballot_sync(0xffffffffu, hit);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const int lj = __shfl_sync(0xffffffffu, li, src);
                const float4 bb = s.lbox[lj];
                const bool mine = open && !(x0 < bb.x || x0 >= bb.z || y0 < bb.y || y0 >= bb.w);
                if (__any_sync(0xffffffffu, mine)) {
                    const int f = s.lid[lj];
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(recs + f));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(recs + f) + 1);
                    if (mine) {
                        const SoftHit h = soft_distance(g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, x0, y0, sentinel);
                        float p, om;
                        soft_prob(h.d2 * zscale, p, om);
                        q = fmaf(p, cc, q);        // 1 - prod(1-p), accurate for small p
                        cc = cc * om;              // prod(1-p), accurate for p near 1
                        c++;
                        if (c >= knum) {           // the K-th accepted face closes the pixel
                            open = false;
                            imidx_img[(size_t)(ty0 + ly) * width + (tx0 + lx)] = -(f + 1);
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, open)) break;
        }
        if (valid && s.cnt[pix] != 255) { s.soft_q[pix] = q; s.soft_c[pix] = cc; s.cnt[pix] = (unsigned char)c; }
    }
    __syncthreads();
    if (tid == 0) s.next_block = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(FWD_THREADS, 2)
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
        mbar_fence_init();
        s.nbig = 0; s.next_block = 0; s.lcount = 0; s.flag = 0; s.sub_uncovered = 0u;
    }
    for (int i = tid; i < TILE * TILE; i += FWD_THREADS) {
        s.zkey[i] = 0ull; s.soft_q[i] = 0.f; s.soft_c[i] = 1.f; s.cnt[i] = 0;
    }
    __syncthreads();
    const float tx_lo = s.xs[0], tx_hi = s.xs[tw - 1];
    const float ty_hi = s.ys[0], ty_lo = s.ys[th - 1];
    const float ex = P.expand_mul;
    uint32_t phase0 = 0, phase1 = 0;

    // whole-image cull: does the union of expanded bboxes touch this tile?  (imgbox holds ordered
    // maxima of (-xmin, -ymin, xmax, ymax), zero-initialised = empty)
    bool touched = false;
    if (fnum > 0) {
        const uint4 ib = P.imgbox[b];
        const float ixmin = -ord2f(ib.x), iymin = -ord2f(ib.y), ixmax = ord2f(ib.z), iymax = ord2f(ib.w);
        touched = (ib.z != 0u) && (ixmin - ex <= tx_hi) && (ixmax + ex > tx_lo) && (iymin - ex <= ty_hi) && (iymax + ex > ty_lo);
    }

    int nbatch = 0;
    if (touched) {
        int pos = 0;
        while (pos < fnum) {
            pos = fill_list(s, bbox, pos, fnum, ex, tx_lo, tx_hi, ty_lo, ty_hi, phase0, phase1);
            if (s.lcount > 0) raster_list(s, recs, tw, th);
            nbatch++;
        }
    }

    // ---- phase C: resolve --------------------------------------------------------------------------
    const float* __restrict__ fattr = P.face_attr + (size_t)f_lo * 3 * D;
    bool any_unc = false;
    for (int it = 0; it < (TILE * TILE) / FWD_THREADS; it++) {
        const int ly = it * (FWD_THREADS / TILE) + tid / TILE, lx = tid % TILE;
        const bool valid = (lx < tw) && (ly < th);
        bool unc = false;
        if (valid) {
            const size_t gp = (size_t)(ty0 + ly) * P.width + (tx0 + lx);
            const unsigned long long key = s.zkey[ly * TILE + lx];
            const size_t px = img_pix + gp;
            if (key != 0ull) {
                const int f = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
                const FaceRec r = recs[f];
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
                    } else {
                        for (int c = 0; c < ch; c++)
                            o[c] = blend(w0, w1, w2, __ldg(a + base + c), __ldg(a + D + base + c), __ldg(a + 2 * D + base + c));
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
                unc = true;
            }
        }
        // which 16x16 sub-tiles still hold uncovered pixels
        const unsigned bal = __ballot_sync(0xffffffffu, unc);
        if (bal) {
            any_unc = true;
            if ((tid & 31) == 0) {
                const int st_row = ly / SUB;
                const int st0 = st_row * (TILE / SUB) + (lx / SUB);       // warp spans 32 columns = 2 sub-tiles
                unsigned m = 0;
                if (bal & 0x0000ffffu) m |= 1u << st0;
                if (bal & 0xffff0000u) m |= 1u << (st0 + 1);
                atomicOr(&s.sub_uncovered, m);
            }
        }
    }
    const int tile_unc = __syncthreads_or(any_unc ? 1 : 0);

    // ---- phase D: soft silhouette ------------------------------------------------------------------
    if (tile_unc && touched && P.knum > 0) {
        const float zscale = (float)P.delta / ((float)P.multiplier * (float)P.multiplier);
        const float sentinel = 4.0f * (float)P.multiplier * (float)P.multiplier;
        if (nbatch == 1) {
            if (s.lcount > 0) soft_list(s, recs, tw, th, P.knum, zscale, sentinel, imidx, P.width, tx0, ty0);
        } else {
            int pos = 0;
            while (pos < fnum) {
                pos = fill_list(s, bbox, pos, fnum, ex, tx_lo, tx_hi, ty_lo, ty_hi, phase0, phase1);
                if (s.lcount > 0) soft_list(s, recs, tw, th, P.knum, zscale, sentinel, imidx, P.width, tx0, ty0);
                // stop early once every uncovered pixel has its K faces
                bool open = false;
                for (int i = tid; i < TILE * TILE; i += FWD_THREADS) open |= ((int)s.cnt[i] < P.knum);
                if (!__syncthreads_or(open ? 1 : 0)) break;
            }
        }
    }
    for (int it = 0; it < (TILE * TILE) / FWD_THREADS; it++) {
        const int ly = it * (FWD_THREADS / TILE) + tid / TILE, lx = tid % TILE;
        if (lx < tw && ly < th && s.cnt[ly * TILE + lx] != 255) {
            const size_t gp = (size_t)(ty0 + ly) * P.width + (tx0 + lx);
            improb[gp] = s.soft_q[ly * TILE + lx];
            imcomp[gp] = s.soft_c[ly * TILE + lx];
        }
    }
}

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
