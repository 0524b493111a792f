// Re-weighted BCE on probabilities -- the consumer of the rendered soft mask on Self6D++'s mask-loss path
// (core/self6dpp/losses/mask_losses.py:63-108 weighted_ex_loss_probs, called on ren_prob at
// core/self6dpp/engine/self_engine_utils.py:541-545):
//     pos = target > 0, neg = target == 0, p = clamp(probs, 1e-7, 1 - 1e-7)
//     loss = sum_pos(-target log p * w) / |pos|  +  sum_neg(-log(1 - p) * w) / |neg|      (a term with an empty set is dropped)
// The reference builds four boolean-indexed temporaries and syncs with the host four times (two `.any()`, two `if num >
// 0`).  Here: one pass that produces the two sums and the two counts (per-thread strided accumulation, fixed shuffle /
// shared-memory trees, per-CTA partials added in CTA order by the CTA that finishes last -- bit-reproducible), and one
// elementwise pass for d loss / d probs.  No host sync.
#include "dibr_internal.h"

namespace dibr {

constexpr int ML_T = 256;

__device__ __forceinline__ float clamp_prob(float p) { return fminf(fmaxf(p, 1e-7f), 1.0f - 1e-7f); }

__global__ void __launch_bounds__(ML_T) mask_loss_forward_kernel(MaskLossParams P)
{
    __shared__ float red[ML_T / 32][4];
    __shared__ int last;
    float s_pos = 0.f, s_neg = 0.f, n_pos = 0.f, n_neg = 0.f;
    for (long long i = (long long)blockIdx.x * ML_T + threadIdx.x; i < P.n; i += (long long)gridDim.x * ML_T) {
        const float t = P.target[i];
        const float w = P.weight ? P.weight[i] : 1.0f;
        const float p = clamp_prob(P.probs[i]);
        if (t > 0.0f) { s_pos += -t * logf(p) * w; n_pos += 1.0f; }
        else if (t == 0.0f) { s_neg += -logf(1.0f - p) * w; n_neg += 1.0f; }
    }
    float v[4] = {s_pos, s_neg, n_pos, n_neg};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ML_T / 32; w++) s += red[w][threadIdx.x];
        P.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (unsigned b = 0; b < gridDim.x; b++) s += __ldcg(P.partial + (size_t)b * 4 + threadIdx.x);     // CTA order: fixed
        red[0][threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float sp = red[0][0], sn = red[0][1], np = red[0][2], nn = red[0][3];
        float loss = 0.f;
        if (np > 0.f) loss += 1.0f / np * sp;
        if (nn > 0.f) loss += 1.0f / nn * sn;
        P.out[0] = loss; P.out[1] = np; P.out[2] = nn;
        *P.ticket = 0u;                          // re-armed for the next call
    }
}

// d loss / d probs (clamp passes the gradient inside [1e-7, 1 - 1e-7] only)
__global__ void __launch_bounds__(ML_T) mask_loss_backward_kernel(MaskLossParams P)
{
    const float np = P.out[1], nn = P.out[2], go = P.grad_out[0];
    for (long long i = (long long)blockIdx.x * ML_T + threadIdx.x; i < P.n; i += (long long)gridDim.x * ML_T) {
        const float t = P.target[i];
        const float w = P.weight ? P.weight[i] : 1.0f;
        const float p = P.probs[i];
        float g = 0.f;
        if (p >= 1e-7f && p <= 1.0f - 1e-7f) {
            if (t > 0.0f) g = -t * w / p / np;
            else if (t == 0.0f) g = w / (1.0f - p) / nn;
        }
        P.grad_probs[i] = g * go;
    }
}

static inline int ml_grid(long long n) {
    const long long want = (n + ML_T * 4 - 1) / (ML_T * 4);
    return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}
int mask_loss_partial_floats(long long n) { return 4 * ml_grid(n); }

int launch_mask_loss_forward(const MaskLossParams& P, cudaStream_t stream)
{
    mask_loss_forward_kernel<<<ml_grid(P.n), ML_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_mask_loss_backward(const MaskLossParams& P, cudaStream_t stream)
{
    if (P.n <= 0) return 0;
    mask_loss_backward_kernel<<<ml_grid(P.n), ML_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr

// ---------------------------------------------------------------------------------------------------------------
// Reduction of the chamfer distances into the depth loss of core/self6dpp/losses/depth_bp_chamfer_loss.py:38-62:
//   per sample  cur = mean(dist1[dist1 < thr]) + mean(dist2[dist2 < thr])   (thr <= 0: all points), NaN when a
//   selection is empty -> the sample is skipped;   loss = sum(cur over the valid samples) / max(#valid, 1)
// One CTA per sample (strided sums, fixed trees), the last CTA adds the samples in order; one elementwise backward.
// ---------------------------------------------------------------------------------------------------------------
namespace dibr {

constexpr int CR_T = 512;

__global__ void __launch_bounds__(CR_T) chamfer_reduce_forward_kernel(ChamferReduceParams P)
{
    __shared__ float red[CR_T / 32][4];
    __shared__ int last;
    const int b = blockIdx.x;
    const int n1 = P.count1 ? min(P.count1[b], P.stride1) : P.stride1, n2 = P.count2 ? min(P.count2[b], P.stride2) : P.stride2;
    const float* d1 = P.dist1 + (size_t)b * P.stride1;
    const float* d2 = P.dist2 + (size_t)b * P.stride2;
    const bool all = !(P.threshold > 0.0f);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < n1; i += CR_T) { const float d = d1[i]; if (all || d < P.threshold) { v[0] += d; v[1] += 1.0f; } }
    for (int i = threadIdx.x; i < n2; i += CR_T) { const float d = d2[i]; if (all || d < P.threshold) { v[2] += d; v[3] += 1.0f; } }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CR_T / 32; w++) s += red[w][threadIdx.x];
        P.stats[(size_t)b * 4 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    float loss = 0.f, nvalid = 0.f;
    for (int s = 0; s < P.batch; s++) {                       // sample order: fixed
        const float s1 = __ldcg(P.stats + (size_t)s * 4), c1 = __ldcg(P.stats + (size_t)s * 4 + 1);
        const float s2 = __ldcg(P.stats + (size_t)s * 4 + 2), c2 = __ldcg(P.stats + (size_t)s * 4 + 3);
        if (c1 > 0.f && c2 > 0.f) { loss += s1 / c1 + s2 / c2; nvalid += 1.0f; }
    }
    P.out[0] = loss / fmaxf(nvalid, 1.0f);
    P.out[1] = nvalid;
    *P.ticket = 0u;
}

__global__ void __launch_bounds__(256) chamfer_reduce_backward_kernel(ChamferReduceParams P)
{
    const int b = blockIdx.y, which = blockIdx.z;
    const int stride = which == 0 ? P.stride1 : P.stride2;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= stride) return;
    const int* cnt = which == 0 ? P.count1 : P.count2;
    const int n = cnt ? min(cnt[b], stride) : stride;
    const float c1 = P.stats[(size_t)b * 4 + 1], c2 = P.stats[(size_t)b * 4 + 3];
    const float d = (which == 0 ? P.dist1 : P.dist2)[(size_t)b * stride + i];
    float g = 0.f;
    if (i < n && c1 > 0.f && c2 > 0.f && (!(P.threshold > 0.0f) || d < P.threshold))
        g = P.grad_out[0] / (which == 0 ? c1 : c2) / fmaxf(P.out[1], 1.0f);
    (which == 0 ? P.grad_dist1 : P.grad_dist2)[(size_t)b * stride + i] = g;
}

int launch_chamfer_reduce_forward(const ChamferReduceParams& P, cudaStream_t stream)
{
    if (P.batch <= 0) return 0;
    chamfer_reduce_forward_kernel<<<P.batch, CR_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_chamfer_reduce_backward(const ChamferReduceParams& P, cudaStream_t stream)
{
    const int smax = P.stride1 > P.stride2 ? P.stride1 : P.stride2;
    if (P.batch <= 0 || smax <= 0) return 0;
    chamfer_reduce_backward_kernel<<<dim3((smax + 255) / 256, P.batch, 2), 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr

// ---------------------------------------------------------------------------------------------------------------
// Soft dice loss on the rendered soft mask: soft_dice_loss (core/self6dpp/losses/mask_losses.py:444-463), the
// MASK_INIT_REN_LOSS_TYPE == "dice" alternative at core/self6dpp/engine/self_engine_utils.py:546-549:
//   per sample  score = 2 (sum p l + smooth) / (sum p + sum l + smooth + eps);   mean: 1 - sum(score) / num,
//   sum: sum(1 - score), none: 1 - score.   One CTA per sample (strided sums, fixed trees) leaves (sum pl, sum p, sum l);
//   the last CTA adds the scores in sample order.  One elementwise backward:
//   d score / d p_i = 2 (l_i D - (I + smooth)) / D^2  with  D = sum p + sum l + smooth + eps.
// ---------------------------------------------------------------------------------------------------------------
namespace dibr {

constexpr int DL_T = 512;

__device__ __forceinline__ float dice_score(const float* st, float smooth, float eps) {
    return 2.0f * (st[0] + smooth) / (st[1] + st[2] + smooth + eps);
}

__global__ void __launch_bounds__(DL_T) dice_loss_forward_kernel(DiceLossParams P)
{
    __shared__ float red[DL_T / 32][3];
    __shared__ int last;
    const int b = blockIdx.x;
    const float* p = P.probs + (size_t)b * P.per;
    const float* l = P.labels + (size_t)b * P.per;
    float v[3] = {0.f, 0.f, 0.f};
    for (long long i = threadIdx.x; i < P.per; i += DL_T) { const float a = p[i], c = l[i]; v[0] += a * c; v[1] += a; v[2] += c; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < DL_T / 32; w++) s += red[w][threadIdx.x];
        P.stats[(size_t)b * 3 + threadIdx.x] = s;
        red[0][threadIdx.x] = s;
    }
    __syncthreads();
    if (P.reduction == 2) {                                   // none: every sample writes its own value
        if (threadIdx.x == 0) P.out[b] = 1.0f - dice_score(red[0], P.smooth, P.eps);
        return;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    float acc = 0.f;
    for (int s = 0; s < P.num; s++) {                         // sample order: fixed
        const float st[3] = {__ldcg(P.stats + (size_t)s * 3), __ldcg(P.stats + (size_t)s * 3 + 1), __ldcg(P.stats + (size_t)s * 3 + 2)};
        const float sc = dice_score(st, P.smooth, P.eps);
        acc += P.reduction == 0 ? sc : 1.0f - sc;
    }
    P.out[0] = P.reduction == 0 ? 1.0f - acc / (float)P.num : acc;
    *P.ticket = 0u;
}

__global__ void __launch_bounds__(256) dice_loss_backward_kernel(DiceLossParams P)
{
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.per) return;
    const float I = P.stats[(size_t)b * 3] + P.smooth, D = P.stats[(size_t)b * 3 + 1] + P.stats[(size_t)b * 3 + 2] + P.smooth + P.eps;
    const float c = P.reduction == 0 ? P.grad_out[0] / (float)P.num : (P.reduction == 1 ? P.grad_out[0] : P.grad_out[b]);
    const float l = P.labels[(size_t)b * P.per + i];
    P.grad_probs[(size_t)b * P.per + i] = -c * (2.0f * (l * D - I) / (D * D));
}

int launch_dice_loss_forward(const DiceLossParams& P, cudaStream_t stream)
{
    if (P.num <= 0) return 0;
    dice_loss_forward_kernel<<<P.num, DL_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_dice_loss_backward(const DiceLossParams& P, cudaStream_t stream)
{
    if (P.num <= 0 || P.per <= 0) return 0;
    dice_loss_backward_kernel<<<dim3((unsigned)((P.per + 255) / 256), P.num), 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr

// ---------------------------------------------------------------------------------------------------------------
// Normal-map loss between the network's normals and the rendered teacher normals: NORMLoss
// (core/self6dpp/losses/vf_norm_loss.py:56-103), applied to the cropped teacher render at
// core/self6dpp/engine/self_engine_utils.py:667-680:
//   a = mask * out, b = mask * gt;  loss = [mean |a - b|]  +  [sum mask * (1 - cos(a, b)) / #(mask != 0)]
// with cos = sum_c (a / max(|a|, 1e-8)) (b / max(|b|, 1e-8)) (F.cosine_similarity over the channel axis).  The reference
// runs ~15 launches and one host sync (.item()); here one reduction launch (fixed trees, last CTA in CTA order) and one
// elementwise backward for d loss / d out.  Planar [n, 3, hw] normals, [n, hw] mask.
// ---------------------------------------------------------------------------------------------------------------
namespace dibr {

constexpr int NL_T = 256;

__global__ void __launch_bounds__(NL_T) norm_loss_forward_kernel(NormLossParams P)
{
    __shared__ float red[NL_T / 32][3];
    __shared__ int last;
    const long long total = (long long)P.n_img * P.hw;
    float v[3] = {0.f, 0.f, 0.f};                             // sum |a - b|, sum m (1 - cos), #(m != 0)
    for (long long i = (long long)blockIdx.x * NL_T + threadIdx.x; i < total; i += (long long)gridDim.x * NL_T) {
        const long long n = i / P.hw, p = i - n * P.hw;
        const float m = P.mask[i];
        const float* o = P.out_norm + n * 3 * P.hw + p;
        const float* g = P.gt_norm + n * 3 * P.hw + p;
        const float a0 = m * o[0], a1 = m * o[P.hw], a2 = m * o[2 * (long long)P.hw];
        const float b0 = m * g[0], b1 = m * g[P.hw], b2 = m * g[2 * (long long)P.hw];
        v[0] += fabsf(a0 - b0) + fabsf(a1 - b1) + fabsf(a2 - b2);
        const float na = fmaxf(sqrtf(a0 * a0 + a1 * a1 + a2 * a2), 1e-8f), nb = fmaxf(sqrtf(b0 * b0 + b1 * b1 + b2 * b2), 1e-8f);
        const float cs = (a0 / na) * (b0 / nb) + (a1 / na) * (b1 / nb) + (a2 / na) * (b2 / nb);
        v[1] += m * (1.0f - cs);
        v[2] += (m != 0.f) ? 1.0f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NL_T / 32; w++) s += red[w][threadIdx.x];
        P.partial[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    float t[3] = {0.f, 0.f, 0.f};
    for (unsigned b = 0; b < gridDim.x; b++)                  // CTA order: fixed
        for (int k = 0; k < 3; k++) t[k] += __ldcg(P.partial + (size_t)b * 3 + k);
    float loss = 0.f;
    if (P.with_l1) loss += t[0] / (3.0f * (float)total);
    if (P.with_cs) loss += t[1] / t[2];
    P.out[0] = loss; P.out[1] = t[2];
    *P.ticket = 0u;
}

__global__ void __launch_bounds__(NL_T) norm_loss_backward_kernel(NormLossParams P)
{
    const long long total = (long long)P.n_img * P.hw;
    const float go = P.grad_out[0], cnt = P.out[1];
    const float k_l1 = P.with_l1 ? go / (3.0f * (float)total) : 0.f;
    for (long long i = (long long)blockIdx.x * NL_T + threadIdx.x; i < total; i += (long long)gridDim.x * NL_T) {
        const long long n = i / P.hw, p = i - n * P.hw;
        const float m = P.mask[i];
        const float* o = P.out_norm + n * 3 * P.hw + p;
        const float* g = P.gt_norm + n * 3 * P.hw + p;
        float a[3], b[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { a[c] = m * o[c * (long long)P.hw]; b[c] = m * g[c * (long long)P.hw]; }
        const float ra = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), rb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        const float na = fmaxf(ra, 1e-8f), nb = fmaxf(rb, 1e-8f);
        const float cs = (a[0] / na) * (b[0] / nb) + (a[1] / na) * (b[1] / nb) + (a[2] / na) * (b[2] / nb);
        const float k_cs = P.with_cs ? -go * m / cnt : 0.f;  // d loss / d cos at this pixel
        float* out = P.grad_out_norm + n * 3 * P.hw + p;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = a[c] - b[c];
            float ga = k_l1 * (d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f));
            // d cos / d a_c = b_c / (|a| |b|) - cos * a_c / |a|^2   (the clamp of a vanishing norm carries no gradient)
            if (ra > 1e-8f) ga += k_cs * ((b[c] / nb) / na - cs * a[c] / (na * na));
            else ga += k_cs * ((b[c] / nb) / na);
            out[c * (long long)P.hw] = ga * m;                // a = mask * out
        }
    }
}

static inline int nl_grid(long long n) {
    const long long want = (n + NL_T * 2 - 1) / (NL_T * 2);
    return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}
int norm_loss_partial_floats(long long pixels) { return 3 * nl_grid(pixels); }

int launch_norm_loss_forward(const NormLossParams& P, cudaStream_t stream)
{
    norm_loss_forward_kernel<<<nl_grid((long long)P.n_img * P.hw), NL_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
int launch_norm_loss_backward(const NormLossParams& P, cudaStream_t stream)
{
    if ((long long)P.n_img * P.hw <= 0) return 0;
    norm_loss_backward_kernel<<<nl_grid((long long)P.n_img * P.hw), NL_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
