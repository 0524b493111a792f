// Face set-up kernels: build the 64 B face records and the per-tile face bitmaps the forward/backward kernels read.
//
//   setup_faces_kernel   operator-seam mode: inputs are the reference's points3d_bxfx9 /
//                        points2d_bxfx6 / normalz_bxfx1; replaces prepare_tfpoints
//                        (/root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:36-70).
//   setup_meshes_kernel  fused mode: object-space vertices + per-instance camera; replaces the
//                        per-sample Python loop of renderer/vcrender_batch.py:49-102 and
//                        renderer/vertex_shaders/perpsective.py:71-111 (view transform, 4x4
//                        projection, divide, per-face gather, face normal, attribute gather with
//                        the ones channel) for the whole ragged batch in ONE launch.  The fp32
//                        operation order is the one frozen in oracle dibr_oracle_project.
#include "dibr_internal.h"

namespace dibr {

__device__ __forceinline__ int image_of_face(int g, int batch, int faces_per_image, const int32_t* __restrict__ off) {
    if (!off) return g / faces_per_image;
    int lo = 0, hi = batch;              // last b with off[b] <= g
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ void write_tables(const SetupParams& P, int gtid) {
    if (gtid < P.width) P.ws.xs[gtid] = pix_x(gtid, P.width, P.multiplier);
    else if (gtid < P.width + P.height) P.ws.ys[gtid - P.width] = pix_y(gtid - P.width, P.height, P.multiplier);
}

// Tile binning: set the face's bit in the bitmap of every 16x16 tile that holds a pixel centre of its EXPANDED bbox
// (exact: the pixel ranges come from first_col_ge / first_row_lt).  Bitmaps instead of lists: OR is order independent,
// so the forward kernel reads the faces of a tile in ascending order without any sort, and the result is the same run
// to run.
__device__ __forceinline__ void bin_face(const SetupParams& P, int g, int b, int e0, int e1, int q0, int q1) {
    if (e1 <= e0 || q1 <= q0) return;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    const int tx0 = e0 / TILE, tx1 = (e1 - 1) / TILE;
    const int ty0 = q0 / TILE, ty1 = (q1 - 1) / TILE;
    const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
    const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
    const int w0 = f_lo >> 5, nw = ((f_hi - 1) >> 5) - w0 + 1;
    uint32_t* img = P.ws.bins + (size_t)tiles_x * tiles_y * ((size_t)w0 + b) + ((g >> 5) - w0);
    const uint32_t bit = 1u << (g & 31);
    if (P.fwd_impl == 2) {                           // the round-1 forward only reads the bitmaps
        for (int ty = ty0; ty <= ty1; ty++)
            for (int tx = tx0; tx <= tx1; tx++) atomicOr(img + (size_t)(ty * tiles_x + tx) * nw, bit);
        return;
    }
    unsigned int* tb = P.ws.tile_blocks + (size_t)b * tiles_x * tiles_y;
    // per-word block masks: byte (word index) of the same layout as the bins, updated with 32-bit atomics
    unsigned int* wm4 = reinterpret_cast<unsigned int*>(P.ws.wordmask);
    const size_t wbase = (size_t)tiles_x * tiles_y * ((size_t)w0 + b) + ((g >> 5) - w0);
    for (int ty = ty0; ty <= ty1; ty++) {
        // 8x4 blocks of the tile the range meets: rows of blocks (4 pixel rows each) x two halves (8 pixel columns each)
        const int rlo = max(q0 - ty * TILE, 0) >> 2, rhi = (min(q1 - ty * TILE, TILE) - 1) >> 2;
        const unsigned rowbits = ((2u << (2 * rhi + 1)) - 1u) & ~((1u << (2 * rlo)) - 1u);         // both halves of block rows rlo..rhi
        for (int tx = tx0; tx <= tx1; tx++) {
            atomicOr(img + (size_t)(ty * tiles_x + tx) * nw, bit);
            const unsigned halves = ((e0 < tx * TILE + 8) ? 0x55u : 0u) | ((e1 > tx * TILE + 8) ? 0xaau : 0u);
            const unsigned m = rowbits & halves;
            if ((__ldcg(tb + ty * tiles_x + tx) & m) != m) atomicOr(tb + ty * tiles_x + tx, m);
            const size_t wi = wbase + (size_t)(ty * tiles_x + tx) * nw;
            atomicOr(wm4 + (wi >> 2), m << ((wi & 3) * 8));
        }
    }
}

__device__ __forceinline__ void store_face(const SetupParams& P, int g, int b, float ax, float ay, float bx, float by,
                                           float cx, float cy, float az, float bz, float cz, float nz, bool active)
{
    if (!active) return;
    FaceRec r;
    r.ax = ax; r.ay = ay; r.bx = bx; r.by = by; r.cx = cx; r.cy = cy;
    r.az = az; r.bz = bz; r.cz = cz; r.nz = nz; r.image = __int_as_float(b); r.pad1 = 0.f;
    r.cols = r.rows = r.ecols = r.erows = 0u;
    // non-finite corners make the face invisible (torch.min/max would propagate the NaN)
    const bool ok = isfinite(ax) && isfinite(ay) && isfinite(bx) && isfinite(by) && isfinite(cx) && isfinite(cy);
    int e0 = 0, e1 = 0, q0 = 0, q1 = 0;
    if (ok) {
        const float xmin = fminf(ax, fminf(bx, cx)), xmax = fmaxf(ax, fmaxf(bx, cx));     // rasterizer.py:49-52
        const float ymin = fminf(ay, fminf(by, cy)), ymax = fmaxf(ay, fmaxf(by, cy));
        const float ex = P.expand_mul;
        const int W = P.width, H = P.height, M = P.multiplier;
        const int c0 = first_col_ge(xmin, W, M), c1 = first_col_ge(xmax, W, M);
        const int r0 = first_row_lt(ymax, H, M), r1 = first_row_lt(ymin, H, M);
        e0 = first_col_ge(xmin - ex, W, M); e1 = first_col_ge(xmax + ex, W, M);            // rasterizer.py:54-57
        q0 = first_row_lt(ymax + ex, H, M); q1 = first_row_lt(ymin - ex, H, M);
        r.cols = (unsigned)c0 | ((unsigned)c1 << 16); r.rows = (unsigned)r0 | ((unsigned)r1 << 16);
        r.ecols = (unsigned)e0 | ((unsigned)e1 << 16); r.erows = (unsigned)q0 | ((unsigned)q1 << 16);
        // front faces with many pixel centres in their bbox are rasterised by the whole grid (coverage kernel, phase 2)
        if (P.fwd_impl != 2 && nz >= 0.0f && (long long)(c1 - c0) * (r1 - r0) > BIG_FACE_PIXELS) {
            const int slot = atomicAdd(P.ws.big_count, 1);
            if (slot < P.ws.big_cap) P.ws.big_list[slot] = g;
        }
    }
    P.ws.recs[g] = r;
    if (P.fwd_impl != 2) P.ws.fbox[g] = make_uint2(r.ecols, r.erows);
    if (ok) bin_face(P, g, b, e0, e1, q0, q1);
}

__global__ void __launch_bounds__(256) setup_faces_kernel(SetupParams P)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    write_tables(P, g);
    const bool active = g < P.total_faces;
    float ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0, az = 0, bz = 0, cz = 0, nz = 0;
    int b = -1;
    if (active) {
        const float m = (float)P.multiplier;
        const float* p2 = P.points2d + (size_t)g * 6;
        ax = __fmul_rn(m, p2[0]); ay = __fmul_rn(m, p2[1]);      // rasterizer.py:46
        bx = __fmul_rn(m, p2[2]); by = __fmul_rn(m, p2[3]);
        cx = __fmul_rn(m, p2[4]); cy = __fmul_rn(m, p2[5]);
        const float* p3 = P.points3d + (size_t)g * 9;
        az = p3[2]; bz = p3[5]; cz = p3[8];
        nz = P.normalz[g];
        b = image_of_face(g, P.batch, P.faces_per_image, P.face_offsets);
    }
    store_face(P, g, b, ax, ay, bx, by, cx, cy, az, bz, cz, nz, active);
}

#ifndef DIBR_SETUP_MIN_CTAS
#define DIBR_SETUP_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(256, DIBR_SETUP_MIN_CTAS) setup_meshes_kernel(SetupParams P)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    write_tables(P, g);
    const bool active = g < P.total_faces;
    float x2[3] = {0, 0, 0}, y2[3] = {0, 0, 0}, zc[3] = {0, 0, 0};
    float nz = 0.f;
    int b = -1;
    // instance lookup once per warp, all lanes at once: lane i looks at instances i, i+32, ...; the instance of the warp's
    // first face is the number of instances that start at or before it, minus one (faces of an instance are
    // contiguous, so lanes then only ever step forward a little)
    int inst0 = 0;
    {
        const int gw = min(blockIdx.x * blockDim.x + (threadIdx.x & ~31), max(P.total_faces - 1, 0));
        int below = 0;
        for (int i0 = 0; i0 < P.num_instances; i0 += 32) {
            const int i = i0 + (threadIdx.x & 31);
            const bool le = (i < P.num_instances) && (__ldg(P.inst_desc + i * INST_STRIDE + I_OUT_FACE_BASE) <= gw);
            below += __popc(__ballot_sync(0xffffffffu, le));
        }
        inst0 = max(below - 1, 0);
    }
    if (active) {
        int inst = inst0;
        while (inst + 1 < P.num_instances && P.inst_desc[(inst + 1) * INST_STRIDE + I_OUT_FACE_BASE] <= g) inst++;
        const int32_t* de = P.inst_desc + inst * INST_STRIDE;
        const int lf = g - de[I_OUT_FACE_BASE];
        b = de[I_IMAGE];
        const int32_t* fv = P.mesh_faces + (size_t)(de[I_MESH_FACE_BASE] + lf) * 3;
        // camera of the instance.  Pose mode: cam_view_R = diag(1,-1,-1) R, cam_view_pos = -(R^T t)
        // (renderer/base.py:169-170) and the 4x4 projection of utils/perspective.py:122-129 straight from R, t, K, in
        // the fp32 operation order frozen in oracle dibr_oracle_camera; the first face of the instance also leaves them
        // in the workspace for the backward (no separate camera kernel).
        float R[9], T[3], Pm[16];
        if (P.pose_R) {
            const float* Rp = P.pose_R + (size_t)de[I_CAM] * 9;
            const float* Tp = P.pose_t + (size_t)de[I_CAM] * 3;
            const float* K = P.pose_K + (size_t)de[I_PROJ] * 9;
            const float w = (float)P.width, h = (float)P.height;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                R[k] = Rp[k]; R[3 + k] = -Rp[3 + k]; R[6 + k] = -Rp[6 + k];
                T[k] = -__fmaf_rn(Rp[6 + k], Tp[2], __fmaf_rn(Rp[3 + k], Tp[1], __fmul_rn(Rp[k], Tp[0])));
            }
#pragma unroll
            for (int i = 0; i < 16; i++) Pm[i] = 0.f;
            Pm[0] = __fdiv_rn(__fmul_rn(2.f, K[0]), w);
            Pm[4] = __fdiv_rn(__fmul_rn(-2.f, K[1]), w);
            Pm[5] = __fdiv_rn(__fmul_rn(2.f, K[4]), h);
            Pm[8] = __fdiv_rn(__fadd_rn(__fmul_rn(-2.f, K[2]), w), w);
            Pm[9] = __fdiv_rn(__fsub_rn(__fmul_rn(2.f, K[5]), h), h);
            Pm[10] = P.q;
            Pm[14] = P.qn;
            Pm[11] = -1.0f;
            if (lf == 0) {
                float* cr = P.ws.cam_rot + (size_t)de[I_CAM] * 9;
                float* cp = P.ws.cam_pos + (size_t)de[I_CAM] * 3;
                float* pm = P.ws.cam_proj + (size_t)de[I_PROJ] * 16;     // instances that share K write the same values
#pragma unroll
                for (int k = 0; k < 9; k++) cr[k] = R[k];
#pragma unroll
                for (int k = 0; k < 3; k++) cp[k] = T[k];
#pragma unroll
                for (int k = 0; k < 16; k++) pm[k] = Pm[k];
            }
        } else {
            const float* Rg = P.cam_rot + (size_t)de[I_CAM] * 9;
            const float* Tg = P.cam_pos + (size_t)de[I_CAM] * 3;
            const float* Pg = P.cam_proj + (size_t)de[I_PROJ] * 16;
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = Rg[k];
#pragma unroll
            for (int k = 0; k < 3; k++) T[k] = Tg[k];
#pragma unroll
            for (int k = 0; k < 16; k++) Pm[k] = Pg[k];
        }
        const float m = (float)P.multiplier;
        const int A = P.vert_attr_dim, D = P.num_attr;
        const bool vec_verts = (P.verts_stride == 4) && ((reinterpret_cast<uintptr_t>(P.verts) & 15) == 0);
        const bool vec_attr = A > 0 && (P.vert_attr_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.vert_attr) & 15) == 0);
        float pc[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int vid = fv[c];
            float vx, vy, vz;
            if (vec_verts) {                                     // rows padded to 16 B: one 128-bit gather
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(P.verts) + (de[I_VERT_BASE] + vid));
                vx = v4.x; vy = v4.y; vz = v4.z;
            } else {
                const float* v = P.verts + (size_t)(de[I_VERT_BASE] + vid) * P.verts_stride;
                vx = v[0]; vy = v[1]; vz = v[2];
            }
            const float d0 = __fsub_rn(vx, T[0]), d1 = __fsub_rn(vy, T[1]), d2 = __fsub_rn(vz, T[2]);
#pragma unroll
            for (int j = 0; j < 3; j++)
                pc[c][j] = __fmaf_rn(R[j * 3 + 2], d2, __fmaf_rn(R[j * 3 + 1], d1, __fmul_rn(R[j * 3 + 0], d0)));
            float clip[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                clip[k] = __fadd_rn(__fmaf_rn(pc[c][2], Pm[8 + k], __fmaf_rn(pc[c][1], Pm[4 + k], __fmul_rn(pc[c][0], Pm[k]))), Pm[12 + k]);
            const float xn = __fdiv_rn(clip[0], clip[3]), yn = __fdiv_rn(clip[1], clip[3]);
            x2[c] = __fmul_rn(m, xn); y2[c] = __fmul_rn(m, yn);
            zc[c] = pc[c][2];
            // per-face corner attributes: [vertex attrs | ones | view depth], 128-bit stores when D % 4 == 0
            float* fa = P.face_attr + ((size_t)g * 3 + c) * D;
            const float* va = P.vert_attr + (size_t)(de[I_ATTR_BASE] + vid) * P.vert_attr_stride;
            float raw[DIBR_MAX_ATTR_INTERNAL];
            if (vec_attr) {                                      // rows padded to a multiple of 16 B
#pragma unroll
                for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d += 4) {
                    if (d < A) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(va + d));
                        raw[d] = t.x; raw[d + 1] = t.y; raw[d + 2] = t.z; raw[d + 3] = t.w;
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) if (d < A) raw[d] = __ldg(va + d);
            }
            float av[DIBR_MAX_ATTR_INTERNAL];
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) {
                float x = 0.f;
                if (d < A) x = raw[d];
                else if (d == A && (P.attr_flags & 1)) x = 1.0f;
                else if (d == A + (P.attr_flags & 1) && (P.attr_flags & 2)) x = -pc[c][2];
                av[d] = x;
            }
            if ((D & 3) == 0) {
#pragma unroll
                for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d += 4)
                    if (d < D) reinterpret_cast<float4*>(fa)[d >> 2] = make_float4(av[d], av[d + 1], av[d + 2], av[d + 3]);
            } else {
#pragma unroll
                for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++)
                    if (d < D) fa[d] = av[d];
            }
        }
        const float e1x = __fsub_rn(pc[1][0], pc[0][0]), e1y = __fsub_rn(pc[1][1], pc[0][1]), e1z = __fsub_rn(pc[1][2], pc[0][2]);
        const float e2x = __fsub_rn(pc[2][0], pc[0][0]), e2y = __fsub_rn(pc[2][1], pc[0][1]), e2z = __fsub_rn(pc[2][2], pc[0][2]);
        const float nx = __fmaf_rn(e1y, e2z, -__fmul_rn(e1z, e2y));
        const float ny = __fmaf_rn(e1z, e2x, -__fmul_rn(e1x, e2z));
        nz = __fmaf_rn(e1x, e2y, -__fmul_rn(e1y, e2x));
        if (P.face_normal) {
            const float len = sqrtf(nx * nx + ny * ny + nz * nz) + 1e-15f;   // utils/utils.py:27-33
            P.face_normal[(size_t)g * 3 + 0] = nx / len;
            P.face_normal[(size_t)g * 3 + 1] = ny / len;
            P.face_normal[(size_t)g * 3 + 2] = nz / len;
        }
    }
    store_face(P, g, b, x2[0], y2[0], x2[1], y2[1], x2[2], y2[2], zc[0], zc[1], zc[2], nz, active);
}

// Work plan of the forward kernel: tiles bucketed by how many faces their bitmap lists (one warp per tile).  The
// forward kernel takes the buckets heaviest first, so the long tiles start early and the short ones fill the tail.
constexpr int PLAN_TILES_PER_CTA = 32;          // one warp per tile, 1024 threads: the bucket counters are bumped once per CTA
__global__ void __launch_bounds__(32 * PLAN_TILES_PER_CTA) plan_tiles_kernel(SetupParams P)
{
    __shared__ int hist[ORDER_BUCKETS], base[ORDER_BUCKETS];
    const int tiles_x = (P.width + TILE - 1) / TILE;
    const int tiles = tiles_x * ((P.height + TILE - 1) / TILE);
    const int ntiles = tiles * P.batch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < ORDER_BUCKETS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x * PLAN_TILES_PER_CTA + warp;
    int kb = -1, pos = 0, id = 0;
    int4 desc = make_int4(0, 0, 0, 0);
    if (t < ntiles) {
        const int b = t / tiles, tl = t - b * tiles;
        const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
        const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
        const int w0 = f_lo >> 5, nw = ((f_hi - 1) >> 5) - w0 + 1;
        const uint32_t* words = P.ws.bins + (size_t)tiles * ((size_t)w0 + b) + (size_t)tl * nw;
        int cost = 0;
        for (int w = lane; w < nw; w += 32) cost += __popc(__ldg(words + w));
        cost = __reduce_add_sync(0xffffffffu, cost);
        kb = min((cost + 31) >> 5, ORDER_BUCKETS - 1);
        const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
        id = (int)(((unsigned)b << 20) | ((unsigned)ty << 10) | (unsigned)tx);                 // forward: unpack_tile
        desc = make_int4(id, f_lo, nw, (int)(words - P.ws.bins));
        if (lane == 0) pos = atomicAdd(&hist[kb], 1);
    }
    __syncthreads();
    if (threadIdx.x < ORDER_BUCKETS && hist[threadIdx.x] > 0) base[threadIdx.x] = atomicAdd(&P.ws.order_cnt[threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
    if (lane == 0 && kb >= 0) {
        P.ws.order_seg[(size_t)kb * ntiles + base[kb] + pos] = id;
        P.ws.tile_desc[(size_t)kb * ntiles + base[kb] + pos] = desc;
    }
    // ---- the CTA that finishes last sums the plan up, so that the forward's CTAs find their tile with one ballot and its
    //      surplus CTAs leave after one load
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&P.ws.order_cnt[PLAN_TICKET], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last && warp == 0) {
        const int n = __ldcg(P.ws.order_cnt + (ORDER_BUCKETS - 1 - lane));
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        P.ws.order_cnt[PLAN_START + lane] = incl - n;
        if (lane == 31) {
            const int touched = incl - n;                       // bucket 0 (lane 31) comes last
            P.ws.order_cnt[PLAN_TOUCHED] = touched;
            P.ws.order_cnt[PLAN_WORK_CTAS] = touched + (n + 7) / 8;
        }
    }
}

static inline int launch_plan(const SetupParams& P, cudaStream_t stream) {
    const int ntiles = ((P.width + TILE - 1) / TILE) * ((P.height + TILE - 1) / TILE) * P.batch;
    plan_tiles_kernel<<<(ntiles + PLAN_TILES_PER_CTA - 1) / PLAN_TILES_PER_CTA, 32 * PLAN_TILES_PER_CTA, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

static inline int setup_grid(const SetupParams& P) {
    const int n = max(P.total_faces, P.width + P.height);
    return (n + 255) / 256;
}

int launch_setup_faces(const SetupParams& P, cudaStream_t stream)
{
    // plan counters, tile bitmaps and the z-buffer are adjacent in the workspace
    cudaError_t e = cudaMemsetAsync(P.ws.order_cnt, 0, P.fwd_impl == 2 ? (size_t)((char*)P.ws.wordmask - (char*)P.ws.order_cnt)
                                                                      : (size_t)((char*)P.ws.zbuf - (char*)P.ws.order_cnt) + P.ws.zbuf_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    setup_faces_kernel<<<setup_grid(P), 256, 0, stream>>>(P);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return launch_plan(P, stream);
}

int launch_setup_meshes(const SetupParams& P, cudaStream_t stream)
{
    // plan counters, tile bitmaps and the z-buffer are adjacent in the workspace
    cudaError_t e = cudaMemsetAsync(P.ws.order_cnt, 0, P.fwd_impl == 2 ? (size_t)((char*)P.ws.wordmask - (char*)P.ws.order_cnt)
                                                                      : (size_t)((char*)P.ws.zbuf - (char*)P.ws.order_cnt) + P.ws.zbuf_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    setup_meshes_kernel<<<setup_grid(P), 256, 0, stream>>>(P);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return launch_plan(P, stream);
}

}  // namespace dibr
