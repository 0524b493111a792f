// Chamfer nearest-neighbour distance, grid-accelerated and still EXACT.
//
// dibr_nnd.cu is the reference's exhaustive search (core/csrc/torch_nndistance/src/nnd_cuda_kernel.cu:8-130) done well;
// it is FP32-issue-bound at n*m pair tests.  The clouds on Self6D++'s path are back-projected depth maps -- surfaces --
// so a uniform grid over the target cloud answers almost every query from a few dozen candidates:
//
//   box      bounding box of every target cloud (ordered-uint atomicMin / atomicMax)
//   count    points per cell of a grid of cubic cells (sqrt(n/8) <= 64 along the longest side), one thread per point
//   scan     exclusive scan of the counts per sample (one CTA, coalesced chunks)
//   fill     counting sort: float4 (x, y, z, original index) written at the cell's cursor
//   query    one thread per query walks cubic shells of cells around its own cell until the best distance found is
//            provably below the distance to everything not yet visited; after NND_RMAX shells it gives up and its
//            WARP scans the whole cloud for it, so far-away queries stay exact and do not become the kernel's tail
//
// The answers are the exhaustive search's, bit for bit: same fp32 distance ((dx*dx + dy*dy) + dz*dz) without
// contraction, and among equal minima the LOWEST original index (= "first minimum in ascending order",
// nnd_cpu.cpp:17), so the order inside a cell -- the only thing the atomic cursor leaves undefined -- cannot matter.
//
// The backward is the same deterministic gather as dibr_nnd.cu but through an inverse index: per target j the list of
// queries k whose nearest neighbour is j (counting sort again), insertion-sorted by k before it is summed, so the
// rounding equals the ascending-k scan of the exhaustive backward (nnd_cpu.cpp:94-130) at O(n + m) instead of O(n m).
#include "dibr_internal.h"

namespace dibr {

constexpr int NND_G = 64;                       // cells per axis at most
constexpr int NND_G3 = NND_G * NND_G * NND_G;
constexpr int NND_RMAX = 4;                     // shells before a query falls back to the exhaustive scan
constexpr int NND_T = 256;
#ifndef DIBR_NND_QL
#define DIBR_NND_QL 1        // measured (tools/nnd_variants.sh): 1 lane 0.50 ms, 2: 0.51, 4: 0.56, 8: 0.70 for the cfg2 batch
#endif
constexpr int NND_QL = DIBR_NND_QL;              // lanes per query: they split every cell range and share the shell walk
static_assert(NND_QL == 1 || NND_QL == 2 || NND_QL == 4 || NND_QL == 8, "a sub-warp group");

// workspace regions:
//   box      [2 dirs][batch][8]   ordered-uint min xyz (words 0..2), max xyz (words 4..6) of the dir's target cloud
//   cells    [2][batch][G3 + 1]   counts -> exclusive starts -> (after the fill) ends
//   sorted   [batch][stride2] for dir 0 then [batch][stride1] for dir 1: (x, y, z, original index)
//   inv_cnt / inv_list            backward: inverse nearest-neighbour index (see below)

__host__ __device__ inline size_t nnd_align(size_t x) { return (x + 255) / 256 * 256; }

size_t nnd_grid_workspace_bytes(int batch, int stride1, int stride2)
{
    size_t n = 0;
    n += nnd_align(sizeof(unsigned int) * 2 * (size_t)batch * 8);
    n += nnd_align(sizeof(int) * 2 * (size_t)batch * (NND_G3 + 1));
    n += nnd_align(sizeof(float4) * (size_t)batch * ((size_t)stride1 + stride2));
    n += nnd_align(sizeof(int) * (size_t)batch * ((size_t)stride1 + stride2 + 2));
    n += nnd_align(sizeof(int) * (size_t)batch * ((size_t)stride1 + stride2));
    return n;
}

struct NndGridPtrs {
    unsigned int* box;
    int* cells;
    float4* sorted[2];      // [dir]: sorted copy of dir's TARGET cloud
    int* inv_cnt[2];        // [dir]: per point of dir's QUERY cloud ... see backward
    int* inv_list[2];
    size_t box_bytes, cells_bytes;
};

static NndGridPtrs nnd_carve(void* base, int batch, int stride1, int stride2)
{
    NndGridPtrs w;
    char* p = (char*)base;
    w.box = (unsigned int*)p; w.box_bytes = sizeof(unsigned int) * 2 * (size_t)batch * 8; p += nnd_align(w.box_bytes);
    w.cells = (int*)p; w.cells_bytes = sizeof(int) * 2 * (size_t)batch * (NND_G3 + 1); p += nnd_align(w.cells_bytes);
    // dir 0: queries = cloud 1, targets = cloud 2 (stride2);  dir 1: targets = cloud 1
    w.sorted[0] = (float4*)p; w.sorted[1] = w.sorted[0] + (size_t)batch * stride2;
    p += nnd_align(sizeof(float4) * (size_t)batch * ((size_t)stride1 + stride2));
    // inverse index of dir 0 (gradient of cloud 1): one counter per cloud-1 point (+1), one list slot per cloud-2 point
    w.inv_cnt[0] = (int*)p; w.inv_cnt[1] = w.inv_cnt[0] + (size_t)batch * (stride1 + 1);
    p += nnd_align(sizeof(int) * (size_t)batch * ((size_t)stride1 + stride2 + 2));
    w.inv_list[0] = (int*)p; w.inv_list[1] = w.inv_list[0] + (size_t)batch * stride2;
    return w;
}

struct CloudView { const float* xyz; int n; };

__device__ __forceinline__ CloudView cloud_of(const NndParams& P, int which, int b) {
    CloudView v;
    const int stride = which == 0 ? P.stride1 : P.stride2;
    const int* cnt = which == 0 ? P.count1 : P.count2;
    v.xyz = (which == 0 ? P.xyz1 : P.xyz2) + (size_t)stride * b * 3;
    v.n = cnt ? min(cnt[b], stride) : stride;
    return v;
}

__device__ __forceinline__ unsigned f2ord_u(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f_u(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ---- box: grid.z = dir, grid.y = sample ------------------------------------------------------------------------
__global__ void __launch_bounds__(NND_T) nnd_box_kernel(NndParams P, unsigned int* box)
{
    const int dir = blockIdx.z, b = blockIdx.y;
    const CloudView T = cloud_of(P, dir == 0 ? 1 : 0, b);
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = blockIdx.x * NND_T + threadIdx.x; k < T.n; k += gridDim.x * NND_T) {
#pragma unroll
        for (int a = 0; a < 3; a++) { const float v = T.xyz[(size_t)k * 3 + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
    unsigned int* o = box + ((size_t)dir * P.batch + b) * 8;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const unsigned l = __reduce_min_sync(0xffffffffu, f2ord_u(lo[a])), h = __reduce_max_sync(0xffffffffu, f2ord_u(hi[a]));
        if ((threadIdx.x & 31) == 0) { atomicMin(o + a, l); atomicMax(o + 4 + a, h); }
    }
}

struct Grid { float lo[3]; float h, inv_h; int d[3]; };

// cubic cells over the target cloud's box; the resolution follows the point count so that a SURFACE sampled by n points
// (what a depth map is) leaves ~8 points in an occupied cell: G = sqrt(n / 8) cells along the longest side
__device__ __forceinline__ Grid grid_of(const unsigned int* box, int n) {
    Grid g;
    float ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) { g.lo[a] = ord2f_u(box[a]); ext = fmaxf(ext, ord2f_u(box[4 + a]) - g.lo[a]); }
    const int G = min(max((int)sqrtf((float)n * 0.125f), 4), NND_G);
    g.h = fmaxf(ext / (float)G * 1.0001f, 1.0e-20f);
    g.inv_h = 1.0f / g.h;
#pragma unroll
    for (int a = 0; a < 3; a++) g.d[a] = min(G, (int)((ord2f_u(box[4 + a]) - g.lo[a]) * g.inv_h) + 1);
    return g;
}
__device__ __forceinline__ int cell_axis(const Grid& g, int a, float v) {
    return min(max((int)floorf((v - g.lo[a]) * g.inv_h), 0), g.d[a] - 1);
}

// ---- count / fill: one thread per target point -----------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(NND_T) nnd_bin_kernel(NndParams P, const unsigned int* box, int* cells, float4* sorted0, float4* sorted1)
{
    const int dir = blockIdx.z, b = blockIdx.y;
    const CloudView T = cloud_of(P, dir == 0 ? 1 : 0, b);
    const int k = blockIdx.x * NND_T + threadIdx.x;
    if (k >= T.n) return;
    const Grid g = grid_of(box + ((size_t)dir * P.batch + b) * 8, T.n);
    const float x = T.xyz[(size_t)k * 3], y = T.xyz[(size_t)k * 3 + 1], z = T.xyz[(size_t)k * 3 + 2];
    const int c = (cell_axis(g, 2, z) * g.d[1] + cell_axis(g, 1, y)) * g.d[0] + cell_axis(g, 0, x);
    int* cs = cells + ((size_t)dir * P.batch + b) * (NND_G3 + 1);
    if (!FILL) { atomicAdd(cs + c, 1); return; }
    const int pos = atomicAdd(cs + c, 1);
    float4* out = (dir == 0 ? sorted0 + (size_t)b * P.stride2 : sorted1 + (size_t)b * P.stride1);
    out[pos] = make_float4(x, y, z, __int_as_float(k));
}

// ---- exclusive scan in place, one CTA of 1024 threads per row, 4096 elements per round with the next round's loads in
//      flight.  mode 0: row = (dir, sample) of the cell counters, length = the cells the sample's grid really has;
//      mode 1 / 2: inverse-index counters of cloud 1 / cloud 2, length = the cloud's point count.
__global__ void __launch_bounds__(1024) nnd_scan_kernel(int* base, size_t row_stride, int mode, NndParams P, const unsigned int* box)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    int* a = base + (size_t)blockIdx.x * row_stride;
    int len;
    if (mode == 0) {
        const int dir = blockIdx.x / P.batch, b = blockIdx.x - dir * P.batch;
        const int n = cloud_of(P, dir == 0 ? 1 : 0, b).n;
        if (n == 0) return;
        const Grid g = grid_of(box + (size_t)blockIdx.x * 8, n);
        len = g.d[0] * g.d[1] * g.d[2];
    } else {
        len = cloud_of(P, mode - 1, blockIdx.x).n;
        if (len == 0) return;
    }
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry = 0;
    int v[4], vn[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (4 * t + k < len) ? a[4 * t + k] : 0;
    for (int c0 = 0; c0 < len; c0 += 4096) {
#pragma unroll
        for (int k = 0; k < 4; k++) vn[k] = (c0 + 4096 + 4 * t + k < len) ? a[c0 + 4096 + 4 * t + k] : 0;      // prefetch
        const int mine = v[0] + v[1] + v[2] + v[3];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wb = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) { const int u = wsum[w]; if (w < warp) wb += u; tot += u; }
        int run = carry + wb + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (c0 + 4 * t + k < len) a[c0 + 4 * t + k] = run; run += v[k]; }
        __syncthreads();
        if (t == 0) carry += tot;
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = vn[k];
    }
}

__device__ __forceinline__ float sqdist_exact(float qx, float qy, float qz, float tx, float ty, float tz) {
    const float x2 = __fsub_rn(tx, qx), y2 = __fsub_rn(ty, qy), z2 = __fsub_rn(tz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(x2, x2), __fmul_rn(y2, y2)), __fmul_rn(z2, z2));
}

#ifdef DIBR_NND_STATS
__device__ unsigned long long g_nnd_stats[16];
extern "C" void dibr_debug_nnd_stats(unsigned long long* out, int reset) { cudaDeviceSynchronize(); cudaMemcpyFromSymbol(out, g_nnd_stats, sizeof(g_nnd_stats)); if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_nnd_stats, z, sizeof(z)); } }
#endif
// ---- query --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NND_T) nnd_query_kernel(NndParams P, const unsigned int* box, const int* cells, const float4* sorted0, const float4* sorted1)
{
    const int dir = blockIdx.z, b = blockIdx.y;
    const CloudView Q = cloud_of(P, dir == 0 ? 0 : 1, b);
    if ((int)(blockIdx.x * (NND_T / NND_QL)) >= Q.n) return;       // whole CTA beyond the cloud
    const int j = (blockIdx.x * NND_T + threadIdx.x) / NND_QL;
    const int lane = threadIdx.x & 31, sub = threadIdx.x % NND_QL;
    const unsigned gmask = (NND_QL == 32) ? 0xffffffffu : (((1u << NND_QL) - 1u) << ((lane / NND_QL) * NND_QL));
    const int m = cloud_of(P, dir == 0 ? 1 : 0, b).n;
    float* dist = (dir == 0 ? P.dist1 + (size_t)P.stride1 * b : P.dist2 + (size_t)P.stride2 * b);
    int* idx = (dir == 0 ? P.idx1 + (size_t)P.stride1 * b : P.idx2 + (size_t)P.stride2 * b);
    if (m == 0) { if (j < Q.n) { dist[j] = 0.f; idx[j] = 0; } return; }      // the exhaustive loop never runs: best = 0, besti = 0
    const bool live = j < Q.n;
    const int jq = live ? j : Q.n - 1;
    const float qx = Q.xyz[(size_t)jq * 3], qy = Q.xyz[(size_t)jq * 3 + 1], qz = Q.xyz[(size_t)jq * 3 + 2];
    const Grid g = grid_of(box + ((size_t)dir * P.batch + b) * 8, m);
    const int* cs = cells + ((size_t)dir * P.batch + b) * (NND_G3 + 1);       // cs[c] = END of cell c after the fill
    const float4* pts = (dir == 0 ? sorted0 + (size_t)b * P.stride2 : sorted1 + (size_t)b * P.stride1);
    const float q[3] = {qx, qy, qz};
    int c[3];
    float scale = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) { c[a] = cell_axis(g, a, q[a]); scale = fmaxf(scale, fmaxf(fabsf(q[a]), fabsf(g.lo[a]) + g.h * g.d[a])); }
    float best = 3.4e38f;
    int bi = 0x7fffffff;
#ifdef DIBR_NND_STATS
    int rr_stat = 0; long long npts_stat = 0;
#endif
    auto scan = [&](int s, int e) {
#ifdef DIBR_NND_STATS
        npts_stat += e - s;
#endif
        for (int k = s + sub; k < e; k += NND_QL) {
            const float4 t = __ldg(pts + k);
            const float d = sqdist_exact(qx, qy, qz, t.x, t.y, t.z);
            const int ti = __float_as_int(t.w);
            if (d < best || (d == best && ti < bi)) { best = d; bi = ti; }
        }
    };
    bool far = false;                            // no answer within NND_RMAX shells: left to the warp (below)
    for (int r = 0; live; r++) {
#ifdef DIBR_NND_STATS
        rr_stat = r;
#endif
        if (r > NND_RMAX) { far = true; break; }
        // shell r of the cube of cells around c
        for (int dz = -r; dz <= r; dz++) {
            const int z = c[2] + dz;
            if (z < 0 || z >= g.d[2]) continue;
            for (int dy = -r; dy <= r; dy++) {
                const int y = c[1] + dy;
                if (y < 0 || y >= g.d[1]) continue;
                const bool face = (dz == -r || dz == r || dy == -r || dy == r);
                const int row = (z * g.d[1] + y) * g.d[0];
                if (face) {
                    // the whole run of cells x in [c0-r, c0+r] is contiguous in memory: one range
                    const int x0 = max(c[0] - r, 0), x1 = min(c[0] + r, g.d[0] - 1);
                    if (x0 <= x1) scan(row + x0 == 0 ? 0 : cs[row + x0 - 1], cs[row + x1]);
                } else {
                    const int xa = c[0] - r, xb = c[0] + r;
                    if (xa >= 0) scan(row + xa == 0 ? 0 : cs[row + xa - 1], cs[row + xa]);
                    if (xb < g.d[0] && xb != xa) scan(cs[row + xb - 1], cs[row + xb]);
                }
            }
        }
        // the lanes of the query agree on the best so far (they walk the shells in lock step, so the group mask is safe)
#pragma unroll
        for (int o = NND_QL / 2; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(gmask, best, o);
            const int oi = __shfl_xor_sync(gmask, bi, o);
            if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
        }
        // everything not visited yet is at least lb away (a side of the cube that already reaches the grid's edge has
        // nothing beyond it); the margins cover the rounding of the cell boundaries and of the distances
        float lb = 3.4e38f;
        bool whole = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (c[a] - r > 0) { whole = false; lb = fminf(lb, q[a] - (g.lo[a] + (float)(c[a] - r) * g.h)); }
            if (c[a] + r < g.d[a] - 1) { whole = false; lb = fminf(lb, (g.lo[a] + (float)(c[a] + r + 1) * g.h) - q[a]); }
        }
        if (whole) break;
        lb = lb * 0.999f - 4.0e-6f * scale;
        if (lb > 0.f && best <= lb * lb) break;
    }
    // queries far from every target point (a part of one cloud the other does not have): the exhaustive scan, done by
    // the whole warp for one such query at a time -- a lone lane walking the cloud would set the kernel's duration
    __syncwarp();
    unsigned todo = __ballot_sync(0xffffffffu, far && sub == 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float fx = __shfl_sync(0xffffffffu, qx, src), fy = __shfl_sync(0xffffffffu, qy, src), fz = __shfl_sync(0xffffffffu, qz, src);
        float wb = 3.4e38f;
        int wi = 0x7fffffff;
        for (int k = lane; k < m; k += 32) {
            const float4 t = __ldg(pts + k);
            const float d = sqdist_exact(fx, fy, fz, t.x, t.y, t.z);
            const int ti = __float_as_int(t.w);
            if (d < wb || (d == wb && ti < wi)) { wb = d; wi = ti; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, wb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (od < wb || (od == wb && oi < wi)) { wb = od; wi = oi; }
        }
        if (lane / NND_QL == src / NND_QL) { best = wb; bi = wi; }
    }
#ifdef DIBR_NND_STATS
    if (live && sub == 0) { atomicAdd(&g_nnd_stats[min(rr_stat, 7)], 1ull); atomicAdd(&g_nnd_stats[8], (unsigned long long)npts_stat); }
#endif
    if (live && sub == 0) {
        // Non-finite input.  The exhaustive search (dibr_nnd.cu, nnd_cpu.cpp:17) starts from target 0 and replaces it only on
        // 'd < best': when d(query, target 0) is NaN nothing ever replaces it, and when no distance is below 3.4e38 the
        // loops above accepted nothing -- in both cases its answer is (d(query, target 0), 0).
        const float* t0 = cloud_of(P, dir == 0 ? 1 : 0, b).xyz;
        const float d0 = sqdist_exact(qx, qy, qz, t0[0], t0[1], t0[2]);
        if (d0 != d0 || bi == 0x7fffffff) { best = d0; bi = 0; }
        dist[j] = best; idx[j] = bi;
    }
}

int launch_nnd_forward_grid(const NndParams& P, void* workspace, cudaStream_t stream)
{
    if (P.batch <= 0 || (P.stride1 <= 0 && P.stride2 <= 0)) return 0;
    const NndGridPtrs w = nnd_carve(workspace, P.batch, P.stride1, P.stride2);
    cudaError_t e = cudaMemsetAsync(w.box, 0xff, w.box_bytes, stream);                  // min slots: all ones
    if (e != cudaSuccess) return (int)e;
    // max slots (words 4..7 of every record) must start at 0: a strided memset
    e = cudaMemset2DAsync(w.box + 4, 8 * sizeof(unsigned int), 0, 4 * sizeof(unsigned int), 2 * (size_t)P.batch, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(w.cells, 0, w.cells_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    const int smax = max(P.stride1, P.stride2);
    const dim3 gpts((smax + NND_T - 1) / NND_T, P.batch, 2);
    nnd_box_kernel<<<dim3(min((smax + NND_T - 1) / NND_T, 16), P.batch, 2), NND_T, 0, stream>>>(P, w.box);
    nnd_bin_kernel<false><<<gpts, NND_T, 0, stream>>>(P, w.box, w.cells, w.sorted[0], w.sorted[1]);
    nnd_scan_kernel<<<2 * P.batch, 1024, 0, stream>>>(w.cells, NND_G3 + 1, 0, P, w.box);
    nnd_bin_kernel<true><<<gpts, NND_T, 0, stream>>>(P, w.box, w.cells, w.sorted[0], w.sorted[1]);
    const dim3 gq((smax * NND_QL + NND_T - 1) / NND_T, P.batch, 2);
    nnd_query_kernel<<<gq, NND_T, 0, stream>>>(P, w.box, w.cells, w.sorted[0], w.sorted[1]);
    return (int)cudaGetLastError();
}

// ---- backward through an inverse index --------------------------------------------------------------------------
// dir 0 produces the gradient of cloud 1: its own term (needs idx1) minus the pulls of every cloud-2 point whose
// nearest neighbour (idx2) is the cloud-1 point.  inverse index of dir 0: per cloud-1 point j the cloud-2 points k with
// idx2[k] == j.
template <int STAGE>
__global__ void __launch_bounds__(NND_T) nnd_inverse_kernel(NndParams P, int* cnt0, int* cnt1, int* list0, int* list1)
{
    const int dir = blockIdx.z, b = blockIdx.y;
    const int other = dir == 0 ? 1 : 0;                       // the cloud whose nearest-neighbour indices point at us
    const CloudView O = cloud_of(P, other, b);
    const int k = blockIdx.x * NND_T + threadIdx.x;
    if (k >= O.n) return;
    const int sA = dir == 0 ? P.stride1 : P.stride2, sB = dir == 0 ? P.stride2 : P.stride1;
    const int nA = cloud_of(P, dir, b).n;
    if (nA == 0) return;
    const int* idxB = (dir == 0 ? P.idx2 : P.idx1) + (size_t)sB * b;
    int* cnt = (dir == 0 ? cnt0 : cnt1) + (size_t)b * (sA + 1);
    const int j = idxB[k];
    if (STAGE == 0) { atomicAdd(cnt + j, 1); return; }
    int* list = (dir == 0 ? list0 : list1) + (size_t)b * sB;
    list[atomicAdd(cnt + j, 1)] = k;                          // after this cnt[j] = END of j's list
}

__global__ void __launch_bounds__(NND_T) nnd_backward_inv_kernel(NndParams P, const int* cnt0, const int* cnt1, const int* list0, const int* list1)
{
    const int dir = blockIdx.z, b = blockIdx.y;
    const int sA = dir == 0 ? P.stride1 : P.stride2, sB = dir == 0 ? P.stride2 : P.stride1;
    if ((int)(blockIdx.x * NND_T) >= sA) return;
    const int j = blockIdx.x * NND_T + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const CloudView A = cloud_of(P, dir, b), Bc = cloud_of(P, dir == 0 ? 1 : 0, b);
    const bool row = j < sA, live = j < A.n;
    const float* gA = (dir == 0 ? P.graddist1 : P.graddist2) + (size_t)sA * b;
    const float* gB = (dir == 0 ? P.graddist2 : P.graddist1) + (size_t)sB * b;
    const int* idxA = (dir == 0 ? P.idx1 : P.idx2) + (size_t)sA * b;
    const int* idxB = (dir == 0 ? P.idx2 : P.idx1) + (size_t)sB * b;
    float ax = 0.f, ay = 0.f, az = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    bool crowd = false;                    // a target that attracts many queries: left to the warp (below)
    auto pull = [&](int k) {
        // explicit roundings (no contraction): the same sum must come out of the per-thread, the warp and the exhaustive path
        const float gk = gB[k] * 2.0f;
        gx = __fsub_rn(gx, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3], ax)));
        gy = __fsub_rn(gy, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3 + 1], ay)));
        gz = __fsub_rn(gz, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3 + 2], az)));
    };
    if (live && Bc.n > 0) {
        ax = A.xyz[(size_t)j * 3]; ay = A.xyz[(size_t)j * 3 + 1]; az = A.xyz[(size_t)j * 3 + 2];
        const int j2 = idxA[j];
        const float g = gA[j] * 2.0f;
        gx = g * (ax - Bc.xyz[(size_t)j2 * 3]); gy = g * (ay - Bc.xyz[(size_t)j2 * 3 + 1]); gz = g * (az - Bc.xyz[(size_t)j2 * 3 + 2]);
        const int* cnt = (dir == 0 ? cnt0 : cnt1) + (size_t)b * (sA + 1);
        const int* list = (dir == 0 ? list0 : list1) + (size_t)b * sB;
        const int s = (j == 0) ? 0 : cnt[j - 1], e = cnt[j];
        if (e - s <= 8) {
            // ascending k, whatever order the cursor produced: repeatedly take the smallest k above the last one taken
            int last = -1;
            for (int t = s; t < e; t++) {
                int kmin = 0x7fffffff;
                for (int u = s; u < e; u++) { const int k = list[u]; if (k > last && k < kmin) kmin = k; }
                last = kmin;
                pull(kmin);
            }
        } else {
            crowd = true;
        }
    }
    // crowds, one target at a time for the whole warp, the sum still in ascending k like the exhaustive gather:
    //   up to 256 queries: every round the lanes look through the target's list for the smallest k above the last one
    //   taken (strided reads + a warp min), the owner's running sums take that query's pull;
    //   beyond: the warp walks the other cloud's index array 32 entries at a time and hands the matches of a round over
    //   in lane order.
    __syncwarp();
    unsigned todo = __ballot_sync(0xffffffffu, crowd);
    const int* cntw = (dir == 0 ? cnt0 : cnt1) + (size_t)b * (sA + 1);
    const int* listw = (dir == 0 ? list0 : list1) + (size_t)b * sB;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int jt = __shfl_sync(0xffffffffu, j, src);
        const float tx = __shfl_sync(0xffffffffu, ax, src), ty = __shfl_sync(0xffffffffu, ay, src), tz = __shfl_sync(0xffffffffu, az, src);
        float sx = __shfl_sync(0xffffffffu, gx, src), sy = __shfl_sync(0xffffffffu, gy, src), sz = __shfl_sync(0xffffffffu, gz, src);
        const int s = (jt == 0) ? 0 : cntw[jt - 1], e = cntw[jt];
        if (e - s <= 256) {
            int last = -1;
            for (int t = s; t < e; t++) {
                int kmin = 0x7fffffff;
                for (int u = s + lane; u < e; u += 32) { const int k = listw[u]; if (k > last && k < kmin) kmin = k; }
                kmin = __reduce_min_sync(0xffffffffu, kmin);
                last = kmin;
                const float gk = gB[kmin] * 2.0f;
                sx = __fsub_rn(sx, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)kmin * 3], tx)));
                sy = __fsub_rn(sy, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)kmin * 3 + 1], ty)));
                sz = __fsub_rn(sz, __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)kmin * 3 + 2], tz)));
            }
        } else {
            for (int k0 = 0; k0 < Bc.n; k0 += 32) {
                const int k = k0 + lane;
                const bool hit = (k < Bc.n) && (idxB[k] == jt);
                unsigned hits = __ballot_sync(0xffffffffu, hit);
                if (hits == 0u) continue;
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (hit) {
                    const float gk = gB[k] * 2.0f;
                    cx = __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3], tx));
                    cy = __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3 + 1], ty));
                    cz = __fmul_rn(gk, __fsub_rn(Bc.xyz[(size_t)k * 3 + 2], tz));
                }
                while (hits) {
                    const int l = __ffs(hits) - 1;
                    hits &= hits - 1;
                    sx = __fsub_rn(sx, __shfl_sync(0xffffffffu, cx, l));
                    sy = __fsub_rn(sy, __shfl_sync(0xffffffffu, cy, l));
                    sz = __fsub_rn(sz, __shfl_sync(0xffffffffu, cz, l));
                }
            }
        }
        if (lane == src) { gx = sx; gy = sy; gz = sz; }
    }
    if (row) {
        float* out = (dir == 0 ? P.gradxyz1 : P.gradxyz2) + ((size_t)sA * b + j) * 3;
        out[0] = live ? gx : 0.f; out[1] = live ? gy : 0.f; out[2] = live ? gz : 0.f;           // padded rows get zeros
    }
}

int launch_nnd_backward_grid(const NndParams& P, void* workspace, cudaStream_t stream)
{
    if (P.batch <= 0 || (P.stride1 <= 0 && P.stride2 <= 0)) return 0;
    const NndGridPtrs w = nnd_carve(workspace, P.batch, P.stride1, P.stride2);
    cudaError_t e = cudaMemsetAsync(w.inv_cnt[0], 0, sizeof(int) * (size_t)P.batch * ((size_t)P.stride1 + P.stride2 + 2), stream);
    if (e != cudaSuccess) return (int)e;
    const int smax = max(P.stride1, P.stride2);
    const dim3 gpts((smax + NND_T - 1) / NND_T, P.batch, 2);
    nnd_inverse_kernel<0><<<gpts, NND_T, 0, stream>>>(P, w.inv_cnt[0], w.inv_cnt[1], w.inv_list[0], w.inv_list[1]);
    // dir 0 counters: [batch][stride1 + 1];  dir 1 counters: [batch][stride2 + 1]
    if (P.stride1 > 0) nnd_scan_kernel<<<P.batch, 1024, 0, stream>>>(w.inv_cnt[0], (size_t)P.stride1 + 1, 1, P, nullptr);
    if (P.stride2 > 0) nnd_scan_kernel<<<P.batch, 1024, 0, stream>>>(w.inv_cnt[1], (size_t)P.stride2 + 1, 2, P, nullptr);
    nnd_inverse_kernel<1><<<gpts, NND_T, 0, stream>>>(P, w.inv_cnt[0], w.inv_cnt[1], w.inv_list[0], w.inv_list[1]);
    nnd_backward_inv_kernel<<<gpts, NND_T, 0, stream>>>(P, w.inv_cnt[0], w.inv_cnt[1], w.inv_list[0], w.inv_list[1]);
    return (int)cudaGetLastError();
}

}  // namespace dibr
