// Normal-map post-processing of Renderer_dibr.render_batch(mode=["norm"])
// (/root/reference/lib/dr_utils/dib_renderer_x/renderer_dibr.py:281-286):
//   out = (n - min) / (||n - min||_2 + 1e-5) * mask   per pixel, min = the batch-global minimum the forward kernel accumulated.
#include "dibr_internal.h"

namespace dibr {

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


// ---- the same map over the tiles of a rasterised pass ---------------------------------------------------------------
// A pixel of a tile no face reaches is uncovered: mask = 0, so its output is 0 whatever the minimum is.  The forward left
// the plan of the pass (tiles bucketed by face count, the empty bucket last), so only the touched tiles (about a quarter
// of a Self6D++ crop batch) are read and normalised; the others are zero-filled, or skipped altogether when the map is
// computed in place over the forward's own zero fill (out == normals).  64 threads per tile, four pixels each, four tiles
// per CTA turn.
struct NormalTilesParams {
    int batch, height, width;
    const int* order_cnt;
    const int* order_seg;
    const float* n;
    const float* mask;
    const unsigned int* min_ordered;
    float* out;
    int vec_ok;
};

__global__ void __launch_bounds__(256) normal_map_tiles_kernel(const NormalTilesParams P)
{
    const float mn = ord2f(*P.min_ordered);
    const int lane = threadIdx.x & 31;
    const int tiles_x = (P.width + TILE - 1) / TILE, tiles_y = (P.height + TILE - 1) / TILE;
    const int ntiles = tiles_x * tiles_y * P.batch;
    const int touched = __ldg(P.order_cnt + PLAN_TOUCHED);
    const bool inplace = (P.out == P.n);
    const int limit = inplace ? touched : ntiles;
    const int start = __ldg(P.order_cnt + PLAN_START + lane);                  // first position of bucket 31 - lane
    const int q = threadIdx.x & 63;                                            // quad of the tile: row q / 4, columns 4 (q % 4) ..
    for (int pos = blockIdx.x * 4 + (threadIdx.x >> 6); pos < limit; pos += gridDim.x * 4) {
        const unsigned le = __ballot_sync(0xffffffffu, start <= pos);
        const int src = 31 - __clz(le);
        const int tile = __ldg(P.order_seg + (size_t)(ORDER_BUCKETS - 1 - src) * ntiles + (pos - __shfl_sync(0xffffffffu, start, src)));
        const int b = (int)((unsigned)tile >> 20), ty = (tile >> 10) & 1023, tx = tile & 1023;
        const int row = ty * TILE + (q >> 2), col = tx * TILE + (q & 3) * 4;
        if (row >= P.height || col >= P.width) continue;
        const size_t pix = ((size_t)b * P.height + row) * P.width + col;
        const bool live = pos < touched;
        if (P.vec_ok && col + 4 <= P.width) {
            float4 ox = make_float4(0.f, 0.f, 0.f, 0.f), oy = ox, oz = ox;
            if (live) {
                const float4* np4 = reinterpret_cast<const float4*>(P.n + pix * 3);
                const float4 x = __ldcs(np4), y = __ldcs(np4 + 1), z = __ldcs(np4 + 2);
                const float4 m = __ldcs(reinterpret_cast<const float4*>(P.mask + pix));
                normal_map_pixel(x.x, x.y, x.z, m.x, mn, ox.x, ox.y, ox.z);
                normal_map_pixel(x.w, y.x, y.y, m.y, mn, ox.w, oy.x, oy.y);
                normal_map_pixel(y.z, y.w, z.x, m.z, mn, oy.z, oy.w, oz.x);
                normal_map_pixel(z.y, z.z, z.w, m.w, mn, oz.y, oz.z, oz.w);
            }
            float4* op4 = reinterpret_cast<float4*>(P.out + pix * 3);
            op4[0] = ox; op4[1] = oy; op4[2] = oz;
        } else {
            for (int k = 0; k < 4 && col + k < P.width; k++) {
                float oa = 0.f, ob = 0.f, oc = 0.f;
                if (live) normal_map_pixel(P.n[(pix + k) * 3], P.n[(pix + k) * 3 + 1], P.n[(pix + k) * 3 + 2], P.mask[pix + k], mn, oa, ob, oc);
                P.out[(pix + k) * 3] = oa; P.out[(pix + k) * 3 + 1] = ob; P.out[(pix + k) * 3 + 2] = oc;
            }
        }
    }
}

int launch_normal_map_tiles(int batch, int height, int width, const int* order_cnt, const int* order_seg, const float* n, const float* mask,
                            const unsigned int* min_ordered, float* out, cudaStream_t stream)
{
    NormalTilesParams P;
    P.batch = batch; P.height = height; P.width = width; P.order_cnt = order_cnt; P.order_seg = order_seg;
    P.n = n; P.mask = mask; P.min_ordered = min_ordered; P.out = out;
    P.vec_ok = ((((uintptr_t)n | (uintptr_t)mask | (uintptr_t)out) & 15) == 0) && (width % 4 == 0);
    const long long ntiles = (long long)((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE) * batch;
    if (ntiles == 0) return 0;
    const int grid = (int)((ntiles + 3) / 4 < 148 * 8 ? (ntiles + 3) / 4 : 148 * 8);
    normal_map_tiles_kernel<<<grid, 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
