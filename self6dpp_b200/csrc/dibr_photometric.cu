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

// powf(t, 2.4) costs ~100 instructions and this path needs 6 of them per pixel (it made the kernel issue-bound at 0.5 TB/s);
// exp2f(2.4 log2f(t)) with the accurate log2f / exp2f is within 1e-6 relative for t in (0.09, 1.06] and a third of the cost.
__device__ __forceinline__ float pow24(float t) { return exp2f(2.4f * log2f(t)); }
__device__ __forceinline__ float srgb_to_linear(float c)
{   // lab.py:43-45
    return c > 0.04045f ? pow24((c + 0.055f) * (1.0f / 1.055f)) : c * (1.0f / 12.92f);
}
__device__ __forceinline__ float lab_f(float n)
{   // lab.py:55-57 (cbrtf: 1 ulp, ~4x cheaper than powf(n, 1/3); cbrtf(0) = 0 like pow)
    return n > 0.008856f ? cbrtf(n) : __fadd_rn(__fmul_rn(7.787f, n), 4.0f / 29.0f);
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
    q.nx = x * (1.0f / 0.95047f); q.ny = y; q.nz = z * (1.0f / 1.08883f);
    q.fx = lab_f(q.nx); q.fy = lab_f(q.ny); q.fz = lab_f(q.nz);
    const float L = __fsub_rn(__fmul_rn(116.0f, q.fy), 16.0f);
    const float A = __fmul_rn(500.0f, __fsub_rn(q.fx, q.fy));
    const float B = __fmul_rn(200.0f, __fsub_rn(q.fy, q.fz));
    q.l = L * (1.0f / 100.0f);                              // lab.py:75-81: (lab - min) / (max - min)
    q.a = __fsub_rn(A, -110.0f) * (1.0f / 220.0f);
    q.b = __fsub_rn(B, -110.0f) * (1.0f / 220.0f);
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
        const float gL = gl * (1.0f / 100.0f), gA = ga * (1.0f / 220.0f), gB = gb * (1.0f / 220.0f);
        const float gfx = 500.0f * gA;
        const float gfy = 116.0f * gL - 500.0f * gA + 200.0f * gB;
        const float gfz = -200.0f * gB;
        // f = where(n > eps, pow(n, 1/3), 7.787 n + 4/29): both branches receive a gradient (one of them zero)
        auto df = [](float gf, float nn) {
            const bool hi = nn > 0.008856f;
            const float c = cbrtf(nn);                        // n^(-2/3) = 1 / cbrt(n)^2; inf at n == 0 -> 0 * inf = NaN like autograd
            const float g_pow = (hi ? gf : 0.f) * ((1.0f / 3.0f) / (c * c));
            const float g_lin = (hi ? 0.f : gf) * 7.787f;
            return g_pow + g_lin;
        };
        const float gx = df(gfx, R.nx) * (1.0f / 0.95047f), gy = df(gfy, R.ny), gz = df(gfz, R.nz) * (1.0f / 1.08883f);
        const float gsr = 0.412453f * gx + 0.212671f * gy + 0.019334f * gz;
        const float gsg = 0.357580f * gx + 0.715160f * gy + 0.119193f * gz;
        const float gsb = 0.180423f * gx + 0.072169f * gy + 0.950227f * gz;
        auto ds = [](float gs, float c) {
            const bool hi = c > 0.04045f;
            const float t = (c + 0.055f) * (1.0f / 1.055f);
            const float g_pow = (hi ? gs : 0.f) * (2.4f * exp2f(1.4f * log2f(t))) * (1.0f / 1.055f);
            const float g_lin = (hi ? 0.f : gs) * (1.0f / 12.92f);
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

// ---------------------------------------------------------------------------------------------------------------
// (2) MS-SSIM -- core/self6dpp/losses/ssim.py:58-160 as instantiated at core/self6dpp/engine/self_engine.py:352
//     (MS_SSIM(data_range=1.0, normalize=True): 11-tap Gaussian, valid convolution, 5 levels, 2x2 average pooling).
//     Per level the reference runs 10 depthwise cuDNN convolutions (TF32 by default on this class of GPU) and ~25
//     elementwise kernels with five (N,C,H,W) temporaries.  Here per level: ONE kernel that stages a 42x42 tile of X and Y
//     in shared memory, runs the separable filter for the five moments (X, Y, XX, YY, XY) in fp32, evaluates cs / ssim per
//     pixel, reduces them per CTA (fixed trees; per-image totals are added in plane / tile order by the combine kernel ->
//     bit-reproducible) and, when a gradient is wanted, stores the three coefficient maps d q / d(mu_y, E[yy], E[xy]) of
//     the one quantity the product uses at that level (cs below the last level, ssim at the last; ssim.py:150-153).
//     Backward per level (coarse to fine): one kernel that runs the transposed separable filter over the three maps and
//     writes  g_y = s * (A + 2 y B + x C) + 0.25 * g_pooled  -- no atomics.
//     The reference's quirk is kept: prod over levels of [cs_l^w_l * ssim_L^w_L] raises the last level's term to the
//     power (levels - 1).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SS_TILE = 32;
constexpr int SS_WIN = 11;
constexpr int SS_IN = SS_TILE + SS_WIN - 1;       // 42

// 2x2 average pooling with zero padding (H % 2, W % 2), count_include_pad (ssim.py:143-145), X and Y in one launch
__global__ void __launch_bounds__(256) ssim_pool_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ px,
                                                        float* __restrict__ py, int planes, int H, int W, int Ho, int Wo)
{
    const int ph = H & 1, pw = W & 1;
    const long long total = (long long)planes * Ho * Wo;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int wo = (int)(i % Wo);
        const long long t = i / Wo;
        const int ho = (int)(t % Ho);
        const long long pl = t / Ho;
        const int r0 = 2 * ho - ph, c0 = 2 * wo - pw;
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int dr = 0; dr < 2; dr++)
#pragma unroll
            for (int dc = 0; dc < 2; dc++) {
                const int r = r0 + dr, c = c0 + dc;
                if (r >= 0 && r < H && c >= 0 && c < W) {
                    const long long o = (pl * H + r) * W + c;
                    sx += x[o]; sy += y[o];
                }
            }
        px[i] = sx * 0.25f; py[i] = sy * 0.25f;
    }
}

__global__ void __launch_bounds__(256) ssim_level_forward_kernel(SsimLevelParams P)
{
    __shared__ float sx[SS_IN][SS_IN + 1];
    __shared__ float sy[SS_IN][SS_IN + 1];
    __shared__ float hz[5][SS_IN][SS_TILE + 1];
    __shared__ float red[8][2];
    // pad = 0: valid windows; pad = 5: zero padding on every side (ssim.py:42-55, use_padding), map of the input's size
    const int Ho = P.H - (SS_WIN - 1) + 2 * P.pad, Wo = P.W - (SS_WIN - 1) + 2 * P.pad;
    const int tiles_x = (Wo + SS_TILE - 1) / SS_TILE;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int plane = blockIdx.y;
    const int r0 = ty * SS_TILE, c0 = tx * SS_TILE;
    const float* X = P.x + (size_t)plane * P.H * P.W;
    const float* Y = P.y + (size_t)plane * P.H * P.W;
    for (int i = threadIdx.x; i < SS_IN * SS_IN; i += 256) {
        const int r = i / SS_IN, c = i - r * SS_IN;
        const int gr = r0 + r - P.pad, gc = c0 + c - P.pad;
        const bool in = gr >= 0 && gc >= 0 && gr < P.H && gc < P.W;
        sx[r][c] = in ? X[(size_t)gr * P.W + gc] : 0.f;
        sy[r][c] = in ? Y[(size_t)gr * P.W + gc] : 0.f;
    }
    __syncthreads();
    // horizontal pass, 4 adjacent outputs per item: 14 + 14 shared-memory loads feed 4 x 5 x 11 FMAs (one output per item
    // needed 22 loads for 55 FMAs and the kernel was bound by shared-memory bandwidth).  Same FMA order per output as before.
    for (int i = threadIdx.x; i < SS_IN * (SS_TILE / 4); i += 256) {
        const int r = i / (SS_TILE / 4), c4 = (i - r * (SS_TILE / 4)) * 4;
        float xv[SS_WIN + 3], yv[SS_WIN + 3];
#pragma unroll
        for (int t = 0; t < SS_WIN + 3; t++) { xv[t] = sx[r][c4 + t]; yv[t] = sy[r][c4 + t]; }
#pragma unroll
        for (int m = 0; m < 5; m++) {
            float src[SS_WIN + 3], a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < SS_WIN + 3; t++)
                src[t] = m == 0 ? xv[t] : m == 1 ? yv[t] : m == 2 ? xv[t] * xv[t] : m == 3 ? yv[t] * yv[t] : xv[t] * yv[t];
#pragma unroll
            for (int k = 0; k < SS_WIN; k++) {
                const float w = P.win[k];
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] = fmaf(w, src[j + k], a[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) hz[m][r][c4 + j] = a[j];
        }
    }
    __syncthreads();
    float s_ssim = 0.f, s_cs = 0.f;
    {
        // vertical pass: a thread owns 4 consecutive rows of one column (14 loads per moment instead of 44)
        const int c = threadIdx.x & (SS_TILE - 1), r4 = (threadIdx.x / SS_TILE) * 4;
        float res[5][4];
#pragma unroll
        for (int m = 0; m < 5; m++) {
            float col[SS_WIN + 3];
#pragma unroll
            for (int t = 0; t < SS_WIN + 3; t++) col[t] = hz[m][r4 + t][c];
#pragma unroll
            for (int j = 0; j < 4; j++) res[m][j] = 0.f;
#pragma unroll
            for (int k = 0; k < SS_WIN; k++) {
                const float w = P.win[k];
#pragma unroll
                for (int j = 0; j < 4; j++) res[m][j] = fmaf(w, col[j + k], res[m][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gr = r0 + r4 + j, gc = c0 + c;
            if (gr >= Ho || gc >= Wo) continue;
            const float mu1 = res[0][j], mu2 = res[1][j], e11 = res[2][j], e22 = res[3][j], e12 = res[4][j];
            const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
            const float s11 = e11 - mu1s, s22 = e22 - mu2s, s12 = e12 - mu12;
            const float dcs = s11 + s22 + P.C2, dl = mu1s + mu2s + P.C1;
            const float cs = (2.f * s12 + P.C2) / dcs;
            const float l = (2.f * mu12 + P.C1) / dl;
            s_ssim += l * cs; s_cs += cs;
            if (P.maps) {
                // d cs / d(mu2, e22, e12)
                float g_mu = (2.f * mu2 * cs - 2.f * mu1) / dcs, g_22 = -cs / dcs, g_12 = 2.f / dcs;
                if (P.use_ssim) {           // last level: d (l * cs)
                    g_mu = l * g_mu + cs * (2.f * (mu1 - l * mu2) / dl);
                    g_22 *= l; g_12 *= l;
                }
                const size_t msz = (size_t)Ho * Wo;
                float* M = P.maps + (size_t)plane * 3 * msz + (size_t)gr * Wo + gc;
                M[0] = g_mu; M[msz] = g_22; M[2 * msz] = g_12;
            }
        }
    }
    float v[2] = {s_ssim, s_cs};
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[w][threadIdx.x];
        P.partial[((size_t)plane * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = s;
    }
}

// per image (one warp each): level means -- lanes take (plane, tile) partials in a fixed interleaved order, then a fixed
// shuffle tree -- the product, and the factors the backward multiplies the maps with
__global__ void __launch_bounds__(128) ssim_combine_kernel(SsimCombineParams P)
{
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= P.n_img) return;
    float ms = 1.0f;
    float csn[SSIM_MAX_LEVELS], ssn_last = 0.f;
    const float half = P.normalize ? 0.5f : 1.0f;
    for (int l = 0; l < P.levels; l++) {
        const int cnt = P.channels * P.tiles[l];                    // this image's partials are contiguous: [channels][tiles][2]
        const float2* q = reinterpret_cast<const float2*>(P.partial + P.partial_off[l]) + (size_t)n * cnt;
        float s_ssim = 0.f, s_cs = 0.f;
        for (int i = lane; i < cnt; i += 32) { const float2 v = q[i]; s_ssim += v.x; s_cs += v.y; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s_ssim += __shfl_xor_sync(0xffffffffu, s_ssim, o); s_cs += __shfl_xor_sync(0xffffffffu, s_cs, o); }
        const float inv = 1.0f / ((float)P.channels * (float)P.map_pixels[l]);
        float ssim = s_ssim * inv, cs = s_cs * inv;
        if (P.normalize) { ssim = (ssim + 1.0f) / 2.0f; cs = (cs + 1.0f) / 2.0f; }       // ssim.py:150-152
        csn[l] = cs;
        if (l == P.levels - 1) ssn_last = ssim;
    }
    if (lane != 0) return;
    const float last_term = powf(ssn_last, P.weights[P.levels - 1]);
    for (int l = 0; l < P.levels - 1; l++) ms *= powf(csn[l], P.weights[l]) * last_term;     // ssim.py:153-156
    P.out[n] = ms;
    if (P.scale) {
        for (int l = 0; l < P.levels - 1; l++)
            P.scale[(size_t)l * P.n_img + n] = ms * P.weights[l] / csn[l] * half / ((float)P.channels * (float)P.map_pixels[l]);
        const int L = P.levels - 1;
        P.scale[(size_t)L * P.n_img + n] = ms * (float)(P.levels - 1) * P.weights[L] / ssn_last * half / ((float)P.channels * (float)P.map_pixels[L]);
    }
}

// g_y(level) = grad_out[n] * scale[level][n] * (G^T mu-map + 2 y G^T e22-map + x G^T e12-map) + 0.25 * g_y(level + 1)[pooled]
__global__ void __launch_bounds__(256) ssim_level_backward_kernel(SsimLevelParams P)
{
    __shared__ float sm[3][SS_IN][SS_IN + 1];
    __shared__ float hz[3][SS_IN][SS_TILE + 1];
    const int Ho = P.H - (SS_WIN - 1) + 2 * P.pad, Wo = P.W - (SS_WIN - 1) + 2 * P.pad;
    const int tiles_x = (P.W + SS_TILE - 1) / SS_TILE;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int plane = blockIdx.y;
    const int n = plane / P.channels;
    const int r0 = ty * SS_TILE, c0 = tx * SS_TILE;
    const size_t msz = (size_t)Ho * Wo;
    const float* M = P.maps + (size_t)plane * 3 * msz;
    // map pixel (u, v) feeds input pixels (u - pad .. u - pad + 10, v - pad .. v - pad + 10): this tile needs
    // u in [r0 + pad - 10, r0 + pad + 31]
    for (int i = threadIdx.x; i < SS_IN * SS_IN; i += 256) {
        const int r = i / SS_IN, c = i - r * SS_IN;
        const int u = r0 + P.pad - (SS_WIN - 1) + r, v = c0 + P.pad - (SS_WIN - 1) + c;
        const bool in = u >= 0 && u < Ho && v >= 0 && v < Wo;
        const size_t o = (size_t)u * Wo + v;
        sm[0][r][c] = in ? M[o] : 0.f; sm[1][r][c] = in ? M[msz + o] : 0.f; sm[2][r][c] = in ? M[2 * msz + o] : 0.f;
    }
    __syncthreads();
    // horizontal: out column c (input pixel c0 + c) = sum_k win[k] * map column (c0 + c - k) = sm[..][c + 10 - k]
    for (int i = threadIdx.x; i < SS_IN * (SS_TILE / 4); i += 256) {       // 4 adjacent outputs per item, as in the forward
        const int r = i / (SS_TILE / 4), c4 = (i - r * (SS_TILE / 4)) * 4;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float src[SS_WIN + 3], a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < SS_WIN + 3; t++) src[t] = sm[m][r][c4 + t];
#pragma unroll
            for (int k = 0; k < SS_WIN; k++) {
                const float w = P.win[k];
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] = fmaf(w, src[j + SS_WIN - 1 - k], a[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) hz[m][r][c4 + j] = a[j];
        }
    }
    __syncthreads();
    const float s = P.scale[n] * P.grad_out[n];
    const int ph = P.H & 1, pw = P.W & 1;
    {
        const int c = threadIdx.x & (SS_TILE - 1), r4 = (threadIdx.x / SS_TILE) * 4;
        float res[3][4];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float col[SS_WIN + 3];
#pragma unroll
            for (int t = 0; t < SS_WIN + 3; t++) col[t] = hz[m][r4 + t][c];
#pragma unroll
            for (int j = 0; j < 4; j++) res[m][j] = 0.f;
#pragma unroll
            for (int k = 0; k < SS_WIN; k++) {
                const float w = P.win[k];
#pragma unroll
                for (int j = 0; j < 4; j++) res[m][j] = fmaf(w, col[j + SS_WIN - 1 - k], res[m][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gr = r0 + r4 + j, gc = c0 + c;
            if (gr >= P.H || gc >= P.W) continue;
            const float A = res[0][j], B = res[1][j], C = res[2][j];
            const size_t o = (size_t)plane * P.H * P.W + (size_t)gr * P.W + gc;
            float g = s * (A + 2.f * P.y[o] * B + P.x[o] * C);
            if (P.grad_coarse) {
                const int pr = (gr + ph) >> 1, pc = (gc + pw) >> 1;
                if (pr < P.Hc && pc < P.Wc) g += 0.25f * P.grad_coarse[(size_t)plane * P.Hc * P.Wc + (size_t)pr * P.Wc + pc];
            }
            P.grad_y[o] = g;
        }
    }
}

int launch_ssim_pool(const float* x, const float* y, float* px, float* py, int planes, int H, int W, int Ho, int Wo, cudaStream_t stream)
{
    const long long total = (long long)planes * Ho * Wo;
    if (total <= 0) return 0;
    ssim_pool_kernel<<<ph_grid(total), 256, 0, stream>>>(x, y, px, py, planes, H, W, Ho, Wo);
    return (int)cudaGetLastError();
}
int ssim_forward_tiles(int H, int W, int pad)
{
    const int Ho = H - (SS_WIN - 1) + 2 * pad, Wo = W - (SS_WIN - 1) + 2 * pad;
    return ((Ho + SS_TILE - 1) / SS_TILE) * ((Wo + SS_TILE - 1) / SS_TILE);
}
int launch_ssim_level_forward(const SsimLevelParams& P, int planes, cudaStream_t stream)
{
    dim3 grid(ssim_forward_tiles(P.H, P.W, P.pad), planes);
    ssim_level_forward_kernel<<<grid, 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_ssim_combine(const SsimCombineParams& P, cudaStream_t stream)
{
    ssim_combine_kernel<<<(P.n_img + 3) / 4, 128, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_ssim_level_backward(const SsimLevelParams& P, int planes, cudaStream_t stream)
{
    dim3 grid(((P.H + SS_TILE - 1) / SS_TILE) * ((P.W + SS_TILE - 1) / SS_TILE), planes);
    ssim_level_backward_kernel<<<grid, 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
