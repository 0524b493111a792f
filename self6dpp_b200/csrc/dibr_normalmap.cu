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

}  // namespace dibr
