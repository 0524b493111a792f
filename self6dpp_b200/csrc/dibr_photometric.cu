// Photometric losses on the rendered colour image.
//
// (1) L1 in normalised CIE-Lab -- core/self6dpp/engine/self_engine_utils.py:745-773 on top of
//     lib/torch_utils/color/lab.py:16-82 (rgb_to_lab, normalize_lab) and lib/torch_utils/color/xyz.py:28-30:
//         lab_x = normalize_lab(rgb_to_lab(x[:, [2, 1, 0]]))            x = gt_img_roi, ren_img_roi  (N,3,H,W, BGR planes)
//         loss  = sum |lab_gt * m - lab_ren * m| / max(1, sum m)        over (a, b) only when LAB_NO_L, m = (N,1,H,W)
//     The reference runs ~60 elementwise torch kernels with (N,3,H,W) temporaries for this.  Here: one pass per direction.
//     Forward: every pixel converts both images in registers and adds its |difference|; the two sums (|diff|, mask) go
//     through fixed shuffle / shared-memory trees, per-CTA partials are added in CTA order by the CTA that finishes last
//     (bit-reproducible, no host sync).  Backward: recomputes the rendered pixel's conversion and applies the chain rule
//     in the order autograd does, INCLUDING its 0 * inf = NaN at exactly-black pixels (pow(0, 1/3) backward) -- the
//     reference produces the same NaN there.
#include "dibr_internal.h"

namespace dibr {

constexpr int PH_T = 256;

struct LabPix { float l, a, b; float fx, fy, fz; float nx, ny, nz; float sr, sg, sb; };

__device__ __forceinline__ float srgb_to_linear(float c)
{   // lab.py:43-45
    return c > 0.04045f ? powf(__fdiv_rn(__fadd_rn(c, 0.055f), 1.055f), 2.4f) : __fdiv_rn(c, 12.92f);
}
__device__ __forceinline__ float lab_f(float n)
{   // lab.py:55-57
    return n > 0.008856f ? powf(n, 1.0f / 3.0f) : __fadd_rn(__fmul_rn(7.787f, n), 4.0f / 29.0f);
}
__device__ __forceinline__ float dot3_rn(float a, float x, float b, float y, float c, float z)
{   // xyz.py:28-30: a*x + b*y + c*z, left to right, no contraction
    return __fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fmul_rn(c, z));
}

// r, g, b in, normalised Lab out (plus the intermediates the backward needs)
__device__ __forceinline__ LabPix rgb_to_lab_norm(float r, float g, float b)
{
    LabPix q;
    q.sr = srgb_to_linear(r); q.sg = srgb_to_linear(g); q.sb = srgb_to_linear(b);
    const float x = dot3_rn(0.412453f, q.sr, 0.357580f, q.sg, 0.180423f, q.sb);
    const float y = dot3_rn(0.212671f, q.sr, 0.715160f, q.sg, 0.072169f, q.sb);
    const float z = dot3_rn(0.019334f, q.sr, 0.119193f, q.sg, 0.950227f, q.sb);
    q.nx = __fdiv_rn(x, 0.95047f); q.ny = __fdiv_rn(y, 1.0f); q.nz = __fdiv_rn(z, 1.08883f);
    q.fx = lab_f(q.nx); q.fy = lab_f(q.ny); q.fz = lab_f(q.nz);
    const float L = __fsub_rn(__fmul_rn(116.0f, q.fy), 16.0f);
    const float A = __fmul_rn(500.0f, __fsub_rn(q.fx, q.fy));
    const float B = __fmul_rn(200.0f, __fsub_rn(q.fy, q.fz));
    q.l = __fdiv_rn(__fsub_rn(L, 0.0f), 100.0f);            // lab.py:75-81: (lab - min) / (max - min)
    q.a = __fdiv_rn(__fsub_rn(A, -110.0f), 220.0f);
    q.b = __fdiv_rn(__fsub_rn(B, -110.0f), 220.0f);
    return q;
}

__device__ __forceinline__ void load_rgb(const float* img, long long n, long long p, long long hw, int bgr, float& r, float& g, float& b)
{
    const float* base = img + n * 3 * hw + p;
    const float c0 = base[0], c1 = base[hw], c2 = base[2 * hw];
    r = bgr ? c2 : c0; g = c1; b = bgr ? c0 : c2;
}

// last-CTA ordered reduction of K per-CTA partials; returns true in the finishing CTA with the totals in tot[]
template <int K>
__device__ __forceinline__ bool ordered_totals(float (&v)[K], float* partial, unsigned int* ticket, float (*red)[K], int* last, float* tot)
{
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PH_T / 32; w++) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * K + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!*last) return false;
    __threadfence();
    if (threadIdx.x < K) {
        float s = 0.f;
        for (unsigned b = 0; b < gridDim.x; b++) s += __ldcg(partial + (size_t)b * K + threadIdx.x);
        red[0][threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) tot[k] = red[0][k];
    return true;
}

__global__ void __launch_bounds__(PH_T) lab_loss_forward_kernel(LabLossParams P)
{
    __shared__ float red[PH_T / 32][2];
    __shared__ int last;
    float v[2] = {0.f, 0.f};
    const long long total = (long long)P.n_img * P.hw;
    for (long long i = (long long)blockIdx.x * PH_T + threadIdx.x; i < total; i += (long long)gridDim.x * PH_T) {
        const long long n = i / P.hw, p = i - n * P.hw;
        const float m = P.mask ? P.mask[i] : 1.0f;
        float r, g, b;
        load_rgb(P.gt, n, p, P.hw, P.bgr, r, g, b);
        const LabPix G = rgb_to_lab_norm(r, g, b);
        load_rgb(P.ren, n, p, P.hw, P.bgr, r, g, b);
        const LabPix R = rgb_to_lab_norm(r, g, b);
        float s = 0.f;
        if (!P.no_l) s = fabsf(__fsub_rn(__fmul_rn(G.l, m), __fmul_rn(R.l, m)));
        s += fabsf(__fsub_rn(__fmul_rn(G.a, m), __fmul_rn(R.a, m)));
        s += fabsf(__fsub_rn(__fmul_rn(G.b, m), __fmul_rn(R.b, m)));
        v[0] += s; v[1] += m;
    }
    float tot[2];
    if (!ordered_totals<2>(v, P.partial, P.ticket, red, &last, tot)) return;
    if (threadIdx.x == 0) {
        const float den = P.mask ? fmaxf(1.0f, tot[1]) : fmaxf(1.0f, (float)total);
        P.out[0] = tot[0] / den; P.out[1] = tot[0]; P.out[2] = den;
        *P.ticket = 0u;
    }
}

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

// d loss / d ren (the gt image is data).  Chain rule in autograd's order; torch.where routes a ZERO into the branch not
// taken and pow's backward multiplies it by exponent * x^(exponent - 1), which is where the reference's NaN at x = 0 comes from.
__global__ void __launch_bounds__(PH_T) lab_loss_backward_kernel(LabLossParams P)
{
    const float scale = P.grad_out[0] / P.out[2];
    const long long total = (long long)P.n_img * P.hw;
    for (long long i = (long long)blockIdx.x * PH_T + threadIdx.x; i < total; i += (long long)gridDim.x * PH_T) {
        const long long n = i / P.hw, p = i - n * P.hw;
        const float m = P.mask ? P.mask[i] : 1.0f;
        float r, g, b;
        load_rgb(P.gt, n, p, P.hw, P.bgr, r, g, b);
        const LabPix G = rgb_to_lab_norm(r, g, b);
        load_rgb(P.ren, n, p, P.hw, P.bgr, r, g, b);
        const LabPix R = rgb_to_lab_norm(r, g, b);
        // d/d(normalised lab of ren) of |G*m - R*m| = -sign(.) * m
        const float gl = P.no_l ? 0.f : -sgnf(__fsub_rn(__fmul_rn(G.l, m), __fmul_rn(R.l, m))) * m * scale;
        const float ga = -sgnf(__fsub_rn(__fmul_rn(G.a, m), __fmul_rn(R.a, m))) * m * scale;
        const float gb = -sgnf(__fsub_rn(__fmul_rn(G.b, m), __fmul_rn(R.b, m))) * m * scale;
        const float gL = gl / 100.0f, gA = ga / 220.0f, gB = gb / 220.0f;
        const float gfx = 500.0f * gA;
        const float gfy = 116.0f * gL - 500.0f * gA + 200.0f * gB;
        const float gfz = -200.0f * gB;
        // f = where(n > eps, pow(n, 1/3), 7.787 n + 4/29): both branches receive a gradient (one of them zero)
        auto df = [](float gf, float nn) {
            const bool hi = nn > 0.008856f;
            const float g_pow = (hi ? gf : 0.f) * ((1.0f / 3.0f) * powf(nn, 1.0f / 3.0f - 1.0f));
            const float g_lin = (hi ? 0.f : gf) * 7.787f;
            return g_pow + g_lin;
        };
        const float gx = df(gfx, R.nx) / 0.95047f, gy = df(gfy, R.ny) / 1.0f, gz = df(gfz, R.nz) / 1.08883f;
        const float gsr = 0.412453f * gx + 0.212671f * gy + 0.019334f * gz;
        const float gsg = 0.357580f * gx + 0.715160f * gy + 0.119193f * gz;
        const float gsb = 0.180423f * gx + 0.072169f * gy + 0.950227f * gz;
        auto ds = [](float gs, float c) {
            const bool hi = c > 0.04045f;
            const float t = (c + 0.055f) / 1.055f;
            const float g_pow = (hi ? gs : 0.f) * (2.4f * powf(t, 1.4f)) / 1.055f;
            const float g_lin = (hi ? 0.f : gs) / 12.92f;
            return g_pow + g_lin;
        };
        float* out = P.grad_ren + n * 3 * P.hw + p;
        const float gr = ds(gsr, r), gg = ds(gsg, g), gbb = ds(gsb, b);
        out[0] = P.bgr ? gbb : gr;
        out[P.hw] = gg;
        out[2 * P.hw] = P.bgr ? gr : gbb;
    }
}

static inline int ph_grid(long long n) {
    const long long want = (n + PH_T - 1) / PH_T;
    return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}
int lab_loss_partial_floats(long long pixels) { return 2 * ph_grid(pixels); }

int launch_lab_loss_forward(const LabLossParams& P, cudaStream_t stream)
{
    lab_loss_forward_kernel<<<ph_grid((long long)P.n_img * P.hw), PH_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_lab_loss_backward(const LabLossParams& P, cudaStream_t stream)
{
    if ((long long)P.n_img * P.hw <= 0) return 0;
    lab_loss_backward_kernel<<<ph_grid((long long)P.n_img * P.hw), PH_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
