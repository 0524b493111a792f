// Set-up kernels: build the 64 B face records, the per-tile face bitmaps and the heaviest-first tile plan the
// forward/backward kernels read.
//
//   setup_faces_kernel   operator-seam mode: inputs are the reference's points3d_bxfx9 /
//                        points2d_bxfx6 / normalz_bxfx1; replaces prepare_tfpoints
//                        (/root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:36-70).
//   setup_meshes_kernel  fused mode: object-space vertices + per-instance camera; replaces the per-sample Python loop
//                        of renderer/vcrender_batch.py:49-102 and renderer/vertex_shaders/perpsective.py:71-111 (view
//                        transform, 4x4 projection, divide, per-face gather, face normal) + prepare_tfpoints for the
//                        whole ragged batch in ONE launch.  Pose mode derives the camera from (R, t, K) first
//                        (renderer/base.py:131-191, utils/perspective.py:95-130).  The fp32 operation order is the one
//                        frozen in oracle dibr_oracle_project / dibr_oracle_camera.  The per-face attribute gather of
//                        vcrender_batch.py:84-88 is NOT materialised: the kernel leaves the three attribute rows of
//                        every face (16 B) and the forward / backward kernels read the vertex table through them.
//                        (The kernel is bound by L1 wavefronts -- distinct 128 B lines per load/store instruction --
//                        not by arithmetic: one 16 B vertex gather per corner plus the redundant transform is cheaper
//                        than a projected-vertex table with 32 B rows, measured in profiles/r02_setup.md.)
//
// Both face kernels end with the tile plan: the binning bumps a per-tile face counter, the CTA that bins the last face
// of an image buckets that image's tiles by cost, and the CTA that plans the last image writes the summary the forward
// kernel's CTAs index with one ballot (no separate plan kernel, no pass over the bitmaps).
#include "dibr_internal.h"

namespace dibr {

__device__ __forceinline__ int image_of_face(int g, int batch, int faces_per_image, const int32_t* __restrict__ off) {
    if (!off) return g / faces_per_image;
    int lo = 0, hi = batch;              // last b with off[b] <= g
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ void write_tables(const SetupParams& P, int gtid) {
    if (gtid == 0 && P.out_min) *P.out_min = 0xffffffffu;          // batch-global minimum of the pass: nothing seen yet
    if (gtid < P.width) P.ws.xs[gtid] = pix_x(gtid, P.width, P.multiplier);
    else if (gtid < P.width + P.height) P.ws.ys[gtid - P.width] = pix_y(gtid - P.width, P.height, P.multiplier);
}

// Tile binning: set the face's bit in the bitmap of every 16x16 tile that holds a pixel centre of its EXPANDED bbox
// (exact: the pixel ranges come from first_col_ge / first_row_lt) and count it in the tile's plan counter.  Bitmaps
// instead of lists: OR is order independent, so the forward kernel reads the faces of a tile in ascending order without
// any sort, and the result is the same run to run.
//
// The 32 faces of a warp are consecutive, so they share ONE bitmap word per tile; where the mesh has locality (neighbouring
// faces land in the same tile) they would send 32 atomics to the same address, which the L2 serialises (a 100k-face mesh:
// set-up 519 us).  So the warp walks its faces' tile rectangles slot by slot (slot j = the j-th tile of every lane's
// rectangle, at most 4 slots), `__match_any_sync` groups the lanes whose slot is the same tile, and one lane per group sends
// a single atomicOr (the group's bits) and a single atomicAdd (its size).  Without locality every lane is its own group
// and the cost is that of the per-face atomics.  (A loop over the DISTINCT tiles of the warp instead of over slots was
// measured 3x slower on meshes without locality: ~80 distinct tiles per warp.)  Every lane of the warp must call.
__device__ __forceinline__ void bin_face_warp(const SetupParams& P, int g, int b, bool ok, int e0, int e1, int q0, int q1) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ok = ok && e1 > e0 && q1 > q0;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles = tiles_x * ((P.height + TILE - 1) / TILE);
    int tx0 = 0, ty0 = 0, ntx = 0, n = 0, w0 = 0, nw = 1;
    if (ok) {
        tx0 = e0 / TILE; ty0 = q0 / TILE;
        ntx = (e1 - 1) / TILE - tx0 + 1;
        n = ntx * ((q1 - 1) / TILE - ty0 + 1);
        const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
        const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
        w0 = f_lo >> 5; nw = ((f_hi - 1) >> 5) - w0 + 1;
    }
    const int nmax = __reduce_max_sync(full, n);
    for (int j = 0; j < min(nmax, 4); j++) {
        int t = -1;                                              // tile of this lane's slot j, unique over the batch with b
        if (j < n) {                                             // j < 4: the row of slot j without an integer division
            const int ry = (j >= ntx ? 1 : 0) + (j >= 2 * ntx ? 1 : 0) + (j >= 3 * ntx ? 1 : 0);
            t = (ty0 + ry) * tiles_x + tx0 + (j - ry * ntx);
        }
        const int key = t < 0 ? -1 : b * tiles + t;
        const unsigned peers = __match_any_sync(full, key);
        if (t >= 0 && lane == __ffs(peers) - 1) {                // bit = lane = g & 31: the CTA starts on a multiple of 32
            atomicOr(P.ws.bins + (size_t)tiles * ((size_t)w0 + b) + (size_t)t * nw + ((g >> 5) - w0), peers);
            atomicAdd(P.ws.tile_count + (size_t)b * tiles + t, __popc(peers));
        }
    }
    // faces that reach more than four tiles are rare: the rest of their rectangle on their own
    for (int j = 4; j < n; j++) {
        const int ry = j / ntx, t = (ty0 + ry) * tiles_x + tx0 + (j - ry * ntx);
        atomicOr(P.ws.bins + (size_t)tiles * ((size_t)w0 + b) + (size_t)t * nw + ((g >> 5) - w0), 1u << (g & 31));
        atomicAdd(P.ws.tile_count + (size_t)b * tiles + t, 1);
    }
}

// Threads per set-up CTA.  157 k faces are ONE wave of work: with 128 threads (56 registers, 10 KB of shared memory) nine CTAs
// fit an SM and the whole grid is resident at once; 256-thread CTAs left 21 of 613 CTAs for a second wave that doubled the
// kernel's time.
#ifndef DIBR_SETUP_THREADS
#define DIBR_SETUP_THREADS 128
#endif
constexpr int SETUP_T = DIBR_SETUP_THREADS;
struct StageSmem {
    float4 rec[SETUP_T * 4];            // 64 B per thread, 16 B chunks swizzled by (thread >> 1) & 3 against bank conflicts
};
__device__ __forceinline__ int rec_slot(int t, int c) { return t * 4 + (c ^ ((t >> 1) & 3)); }

__device__ __forceinline__ void store_face(const SetupParams& P, StageSmem& st, int g, int b, float ax, float ay, float bx, float by,
                                           float cx, float cy, float az, float bz, float cz, float nz, bool active)
{
    // non-finite corners make the face invisible (torch.min/max would propagate the NaN)
    const bool ok = active && isfinite(ax) && isfinite(ay) && isfinite(bx) && isfinite(by) && isfinite(cx) && isfinite(cy);
    int c0 = 0, c1 = 0, r0 = 0, r1 = 0, e0 = 0, e1 = 0, q0 = 0, q1 = 0;
    if (ok) {
        const float xmin = fminf(ax, fminf(bx, cx)), xmax = fmaxf(ax, fmaxf(bx, cx));     // rasterizer.py:49-52
        const float ymin = fminf(ay, fminf(by, cy)), ymax = fmaxf(ay, fmaxf(by, cy));
        const float ex = P.expand_mul;
        const int W = P.width, H = P.height, M = P.multiplier;
        c0 = first_col_ge(xmin, W, M); c1 = first_col_ge(xmax, W, M);
        r0 = first_row_lt(ymax, H, M); r1 = first_row_lt(ymin, H, M);
        e0 = first_col_ge(xmin - ex, W, M); e1 = first_col_ge(xmax + ex, W, M);            // rasterizer.py:54-57
        q0 = first_row_lt(ymax + ex, H, M); q1 = first_row_lt(ymin - ex, H, M);
    }
    const int t = threadIdx.x;
    st.rec[rec_slot(t, 0)] = make_float4(ax, ay, bx, by);
    st.rec[rec_slot(t, 1)] = make_float4(cx, cy, az, bz);
    // (the face's id inside its image rides along: the backward needs it against imidx and would otherwise chain a load of face_offsets[b])
    const int f_lo = (active && b >= 0) ? (P.face_offsets ? __ldg(P.face_offsets + b) : b * P.faces_per_image) : 0;
    st.rec[rec_slot(t, 2)] = make_float4(cz, nz, __int_as_float(b), __int_as_float(g - f_lo));
    st.rec[rec_slot(t, 3)] = make_float4(__uint_as_float((unsigned)c0 | ((unsigned)c1 << 16)), __uint_as_float((unsigned)r0 | ((unsigned)r1 << 16)),
                                         __uint_as_float((unsigned)e0 | ((unsigned)e1 << 16)), __uint_as_float((unsigned)q0 | ((unsigned)q1 << 16)));
    bin_face_warp(P, g, b, ok, e0, e1, q0, q1);
    __syncthreads();
    // the CTA's records are contiguous in global memory: 16 B chunks, four per thread, fully coalesced
    const int g0 = blockIdx.x * blockDim.x;
    const int nrec = min((int)blockDim.x, P.total_faces - g0);
    float4* out = reinterpret_cast<float4*>(P.ws.recs + g0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = k * SETUP_T + t;
        if (i < nrec * 4) out[i] = st.rec[rec_slot(i >> 2, i & 3)];
    }
}

// ---- tile plan -----------------------------------------------------------------------------------------------------
// tile ids of the plan: image | tile row | tile column (forward: unpack_tile)
__device__ __forceinline__ int pack_tile(int b, int ty, int tx) { return (int)(((unsigned)b << 20) | ((unsigned)ty << 10) | (unsigned)tx); }

struct PlanSmem {
    int hist[ORDER_BUCKETS], base[ORDER_BUCKETS], fill[ORDER_BUCKETS];
    int fin[SETUP_T];       // images this CTA completed
    int nfin, last;
};

// The whole CTA buckets the tiles of image b by cost (faces listed in the tile's bitmap, 32 per bucket step) and appends
// them to the global buckets.  The order inside a bucket is arbitrary; it only decides which CTA of the forward grid gets
// which tile.
__device__ void plan_image(const SetupParams& P, PlanSmem& s, int b)
{
    const int tiles_x = (P.width + TILE - 1) / TILE;
    const int tiles = tiles_x * ((P.height + TILE - 1) / TILE);
    const int ntiles = tiles * P.batch;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int* __restrict__ cnt = P.ws.tile_count + (size_t)b * tiles;
    const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
    const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
    const int w0 = f_lo >> 5, nw = ((f_hi - 1) >> 5) - w0 + 1;          // bitmap words per tile (bin_face)
    if (tid < ORDER_BUCKETS) { s.hist[tid] = 0; s.fill[tid] = 0; }
    __syncthreads();
    // most tiles of an image share a bucket (the empty one): the lanes of a warp that agree on the bucket send ONE shared-memory
    // atomic between them (1,200 tiles of a 480x640 image on one counter cost 18 us otherwise)
    const int lane = tid & 31;
    for (int t0 = 0; t0 < tiles; t0 += nthr) {
        const int t = t0 + tid;
        const int kb = t < tiles ? min((__ldcg(cnt + t) + 31) >> 5, ORDER_BUCKETS - 1) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, kb);
        if (kb >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s.hist[kb], __popc(peers));
    }
    __syncthreads();
    if (tid < ORDER_BUCKETS) s.base[tid] = s.hist[tid] > 0 ? atomicAdd(&P.ws.order_cnt[tid], s.hist[tid]) : 0;
    __syncthreads();
    for (int t0 = 0; t0 < tiles; t0 += nthr) {
        const int t = t0 + tid;
        const int kb = t < tiles ? min((__ldcg(cnt + t) + 31) >> 5, ORDER_BUCKETS - 1) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, kb);
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (kb >= 0 && lane == leader) pos = atomicAdd(&s.fill[kb], __popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(peers & ((1u << lane) - 1u));
        if (kb < 0) continue;
        pos += s.base[kb];
        const int ty = t / tiles_x;
        const int id = pack_tile(b, ty, t - ty * tiles_x);
        P.ws.order_seg[(size_t)kb * ntiles + pos] = id;
        P.ws.order_desc[(size_t)kb * ntiles + pos] = make_int4(id, f_lo, f_hi, (int)((size_t)tiles * ((size_t)w0 + b) + (size_t)t * nw));
    }
    __syncthreads();
}

// End of a face kernel.  `g`: this thread's face, `b`: its image (when active).  Every thread of the CTA must call.
__device__ void plan_epilogue(const SetupParams& P, int g, int b, bool active)
{
    __shared__ PlanSmem s;
    const int tid = threadIdx.x;
    if (tid == 0) { s.nfin = 0; s.last = 0; }
    __syncthreads();
    // the CTA's binning atomics before the image counters.  One thread fences for the CTA (fences are cumulative over
    // the barrier): a gpu-scope fence also drops the SM's L1, which the other CTAs of the SM are gathering through.
    if (tid == 0) __threadfence();
    __syncthreads();
    if (active) {
        const int f_lo = P.face_offsets ? P.face_offsets[b] : b * P.faces_per_image;
        const int f_hi = P.face_offsets ? P.face_offsets[b + 1] : f_lo + P.faces_per_image;
        // the last face of image b inside this CTA reports how many of the image's faces the CTA has binned
        if (tid == (int)blockDim.x - 1 || g + 1 == f_hi) {
            const int first = max((int)(blockIdx.x * blockDim.x), f_lo);
            const int n = g - first + 1;
            if (atomicAdd(&P.ws.img_done[b], n) + n == f_hi - f_lo) s.fin[atomicAdd(&s.nfin, 1)] = b;
        }
    }
    __syncthreads();
    int planned = 0;
    // images without faces are nobody's: the first CTA plans them (all their tiles go to bucket 0)
    if (blockIdx.x == 0) {
        for (int i = 0; i < P.batch; i++) {
            const int lo = P.face_offsets ? P.face_offsets[i] : i * P.faces_per_image;
            const int hi = P.face_offsets ? P.face_offsets[i + 1] : lo + P.faces_per_image;
            if (hi <= lo) { plan_image(P, s, i); planned++; }
        }
    }
    const int nfin = s.nfin;
    for (int i = 0; i < nfin; i++) { plan_image(P, s, s.fin[i]); planned++; }
    if (planned == 0) return;
    // ---- the CTA that plans the last image sums the plan up, so that the forward's CTAs find their tile with one
    //      ballot and its surplus CTAs leave after one load
    __syncthreads();
    if (tid == 0) { __threadfence(); s.last = (atomicAdd(&P.ws.order_cnt[PLAN_TICKET], planned) + planned == P.batch); }
    __syncthreads();
    if (s.last && tid < 32) {
        const int lane = tid;
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

// faces in use: with face_offsets the count lives on the device (P.total_faces is then the CAPACITY of the face arrays, which
// keeps the workspace layout and the grid the same from step to step -- a captured CUDA graph stays valid when the batch
// composition changes)
__device__ __forceinline__ int faces_in_use(const SetupParams& P) {
    return P.face_offsets ? min(__ldg(P.face_offsets + P.batch), P.total_faces) : P.total_faces;
}

__global__ void __launch_bounds__(SETUP_T) setup_faces_kernel(SetupParams P)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    write_tables(P, g);
    const bool active = g < faces_in_use(P);
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
    __shared__ StageSmem st;
    store_face(P, st, g, b, ax, ay, bx, by, cx, cy, az, bz, cz, nz, active);
    plan_epilogue(P, g, b, active);
}

// ---- fused mode: one thread per face of the ragged batch ----------------------------------------------------------
__global__ void __launch_bounds__(SETUP_T) setup_meshes_kernel(SetupParams P)
{
    __shared__ StageSmem st;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    write_tables(P, g);
    const int nfaces = faces_in_use(P);
    const bool active = g < nfaces;
    float x2[3] = {0, 0, 0}, y2[3] = {0, 0, 0}, zc[3] = {0, 0, 0};
    float nz = 0.f;
    int b = -1;
    // instance lookup once per warp, all lanes at once: lane i looks at instances i, i+32, ...; the instance of the warp's
    // first face is the number of instances that start at or before it, minus one (faces of an instance are
    // contiguous, so lanes then only ever step forward a little)
    int inst0 = 0;
    {
        const int gw = min(blockIdx.x * blockDim.x + (threadIdx.x & ~31), max(nfaces - 1, 0));
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
        const bool vec_verts = (P.verts_stride == 4) && ((reinterpret_cast<uintptr_t>(P.verts) & 15) == 0);
        float pc[3][3];
        // rows of the three corners in the vertex-attribute table: the forward / backward gather through them
        P.ws.fvid[g] = make_int4(de[I_ATTR_BASE] + fv[0], de[I_ATTR_BASE] + fv[1], de[I_ATTR_BASE] + fv[2], 0);
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
    store_face(P, st, g, b, x2[0], y2[0], x2[1], y2[1], x2[2], y2[2], zc[0], zc[1], zc[2], nz, active);
    plan_epilogue(P, g, b, active);
}


static inline int setup_grid(const SetupParams& P) {
    const int n = max(P.total_faces, P.width + P.height);
    return (n + SETUP_T - 1) / SETUP_T;
}

// plan counters, per-tile face counters, per-image progress counters, the tile bitmaps and the backward's list counters /
// face flags are adjacent in the workspace: one memset
static inline cudaError_t clear_plan(const SetupParams& P, cudaStream_t stream) {
    return cudaMemsetAsync(P.ws.order_cnt, 0, P.ws.clear_bytes, stream);
}

int launch_setup_faces(const SetupParams& P, cudaStream_t stream)
{
    cudaError_t e = clear_plan(P, stream);
    if (e != cudaSuccess) return (int)e;
    setup_faces_kernel<<<setup_grid(P), SETUP_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

int launch_setup_meshes(const SetupParams& P, cudaStream_t stream)
{
    cudaError_t e = clear_plan(P, stream);
    if (e != cudaSuccess) return (int)e;
    setup_meshes_kernel<<<setup_grid(P), SETUP_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
