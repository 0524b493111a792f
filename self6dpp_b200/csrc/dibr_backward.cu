// Backward pass of the B200 DIB-R rasterizer -- deterministic, no global atomics.
//
// The reference (kaolin v0.1 dr_cuda_backward_color_batch / dr_cuda_backward_prob_batch, called at
// /root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:249-269) runs one thread per
// PIXEL and scatters into its face with fp32 atomicAdd (9*D + up to 4*K atomics per pixel, order
// undefined).  Here the loop is turned inside out: the forward kernel leaves two work lists -- faces
// that won a pixel, faces that entered a soft-silhouette product -- and a group of lanes owns one
// listed FACE: it walks the pixel centres inside its bbox (colour part, 4 lanes) or expanded bbox
// (soft part, 16 lanes), picks up the pixels that belong to it (imidx == face+1, or uncovered with
// the face among the first K), accumulates in registers in a fixed order and reduces with a fixed
// shuffle tree.  The order of the work lists is arbitrary but no result depends on it: every face is
// reduced on its own and written to its own slot, so the gradients are bit-reproducible run to run.
//
// Algebra used (see DESIGN.md "Backward"): with acc[i][d] = sum_pix w_i * dL/dI_d (which IS
// dL/dattr, rasterizer.py:278-291 'dldc'),  A_i = sum_d (c1-c0)_d acc[i][d],
// B_i = sum_d (c2-c0)_d acc[i][d], the reference's per-pixel coordinate gradient sums to
//   dL/dP_i = multiplier / k3 * ( -q A_i + p B_i ,  n A_i - m B_i ),
// so the colour part needs no per-pixel coordinate work at all.
#include "dibr_internal.h"

namespace dibr {

#ifndef DIBR_COLOR_LANES
#define DIBR_COLOR_LANES 4        // lanes per face (rows of its bbox per turn): 4 measured best (8: +6 us, 2: +2 us, 16: +16 us)
#endif
constexpr int GRP = DIBR_COLOR_LANES;     // lanes per face in the colour part

// ---- colour part: one 8-lane group per face that WON at least one pixel (work list written by the forward) ------
template <int DMAX>
__device__ __forceinline__ void backward_color_body(const BwdParams& P, int bid)
{
    const int tid = threadIdx.x;
    const int gl = tid & (GRP - 1);
    const int gi = bid * (256 / GRP) + (tid / GRP);
    const int nlist = P.list_counts[0];
    if (bid * (256 / GRP) >= nlist) return;          // whole CTA beyond the list
    const bool active = gi < nlist;
    const int D = P.num_attr;
    const int W = P.width, H = P.height;
    const int shift = (tid & 31) & ~(GRP - 1);                              // first lane of the group inside its warp
    const unsigned full = (GRP == 32) ? 0xffffffffu : (((1u << GRP) - 1u) << shift);

    float acc[3 * DMAX];
#pragma unroll
    for (int i = 0; i < 3 * DMAX; i++) acc[i] = 0.f;
    FaceRec rec;
    int g = 0;
    if (active) {
        g = P.color_list[gi];
        rec = P.recs[g];
        const int b = __float_as_int(rec.image);
        const int f = __float_as_int(rec.local_id);
        const size_t img = (size_t)b * H * W;
        const int32_t* __restrict__ idx = P.imidx + img;
        // pixel centres inside the face's bbox: the ranges the set-up kernel left in the record
        const int c0 = (int)(rec.cols & 0xffffu), c1 = (int)(rec.cols >> 16);
        const int r0 = (int)(rec.rows & 0xffffu), r1 = (int)(rec.rows >> 16);
        if (c1 > c0 && r1 > r0 && P.any_grad_im) {
            const FaceK fk = make_facek(rec);
            // Lane = pixel row of the bbox: a mask of the pixels this face won (about a fifth of the bbox), then the
            // (row, column) pairs of the group are numbered by a prefix sum and dealt out evenly, one per lane and turn:
            // the per-pixel work (weights, gradient loads, 3 D fused multiply-adds) runs with every lane busy.  Which lane
            // adds which pixel depends on the masks alone and the lanes are combined by a fixed tree: bit-reproducible.
            for (int rb = r0; rb < r1; rb += GRP) {
                const int r = rb + gl;
                for (int cbase = c0; cbase < c1; cbase += 32) {
                    unsigned m = 0u;
                    if (r < r1) {
                        const int32_t* __restrict__ ir = idx + (size_t)r * W + cbase;
                        const int ncol = min(32, c1 - cbase);
                        for (int k = 0; k < ncol; k++) m |= (ir[k] == f + 1) ? (1u << k) : 0u;
                    }
                    const int cnt = __popc(m);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < GRP; o <<= 1) {
                        const int t = __shfl_up_sync(full, incl, o, GRP);
                        if (gl >= o) incl += t;
                    }
                    const int total = __shfl_sync(full, incl, GRP - 1, GRP);
                    for (int jb = 0; jb < total; jb += GRP) {
                        const int jq = jb + gl;
                        int src = 0;                            // lanes whose running count is <= jq: the lane that holds pair jq
#pragma unroll
                        for (int st = GRP / 2; st >= 1; st >>= 1) {
                            const int v = __shfl_sync(full, incl, min(src + st - 1, GRP - 1), GRP);
                            if (v <= jq) src += st;
                        }
                        src = min(src, GRP - 1);
                        const unsigned ms = __shfl_sync(full, m, src, GRP);
                        const int before = __shfl_sync(full, incl - cnt, src, GRP);
                        if (jq < total) {
                            const int rr = rb + src, c = cbase + (int)__fns(ms, 0u, jq - before + 1);
                            const size_t pix = (size_t)rr * W + c;
                            float w0, w1, w2;
                            bary(fk, P.xs[c], P.ys[rr], w0, w1, w2);
#pragma unroll
                            for (int d = 0; d < DMAX; d++) {
                                if (d < D && P.chan_grad[d]) {
                                    const float gv = __ldg(P.chan_grad[d] + (img + pix) * (size_t)P.chan_stride[d]);
                                    acc[0 * DMAX + d] = fmaf(gv, w0, acc[0 * DMAX + d]);
                                    acc[1 * DMAX + d] = fmaf(gv, w1, acc[1 * DMAX + d]);
                                    acc[2 * DMAX + d] = fmaf(gv, w2, acc[2 * DMAX + d]);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    // fixed-tree reduction over the lanes of the face (channels without an upstream gradient stay zero: skipped)
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
        if (d < D && P.chan_grad[d]) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float v = acc[i * DMAX + d];
#pragma unroll
                for (int o = GRP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o, GRP);
                acc[i * DMAX + d] = v;
            }
        }
    }
    if (!active || gl != 0) return;
    // dL/dattr is acc itself; dL/dP follows from it (see the header comment)
    const FaceK fk = make_facek(rec);
    // corner attributes: seam mode from face_attr; fused mode [row of the vertex table | 1 | -view z] through the face's row ids
    const bool fusedm = P.va.fvid != nullptr;
    const float* __restrict__ a = fusedm ? nullptr : P.face_attr + (size_t)g * 3 * D;
    int4 vid = make_int4(0, 0, 0, 0);
    if (fusedm) vid = __ldg(P.va.fvid + g);
    const float* __restrict__ a0 = P.va.table + (size_t)vid.x * P.va.stride;
    const float* __restrict__ a1 = P.va.table + (size_t)vid.y * P.va.stride;
    const float* __restrict__ a2 = P.va.table + (size_t)vid.z * P.va.stride;
    const int nA = P.va.dim, ones = P.va.flags & 1, dep = (P.va.flags >> 1) & 1;
    float A[3] = {0.f, 0.f, 0.f}, Bv[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
        if (d < D && P.chan_grad[d]) {                  // (a channel without an upstream gradient has acc == 0: it adds nothing)
            float c0v, e1, e2;
            if (!fusedm) { c0v = a[d]; e1 = a[D + d] - c0v; e2 = a[2 * D + d] - c0v; }
            else if (d < nA) { c0v = __ldg(a0 + d); e1 = __ldg(a1 + d) - c0v; e2 = __ldg(a2 + d) - c0v; }
            else if (dep && d == nA + ones) { c0v = -rec.az; e1 = -rec.bz - c0v; e2 = -rec.cz - c0v; }
            else { c0v = 1.0f; e1 = 0.f; e2 = 0.f; }        // the ones channel: constant over the face
#pragma unroll
            for (int i = 0; i < 3; i++) {
                A[i] = fmaf(e1, acc[i * DMAX + d], A[i]);
                Bv[i] = fmaf(e2, acc[i * DMAX + d], Bv[i]);
            }
        }
    }
    // multiplier * k3 / (k3^2 + eps): the reference's multiplier * g / (k3*k3 + eps) times the k3 the
    // un-normalised dw terms carry
    const float inv = (float)P.multiplier * fk.k3 / (fk.k3 * fk.k3 + 1e-15f);
    float* __restrict__ gpo = P.grad_points2d + (size_t)g * 6;
    float* __restrict__ gao = P.grad_face_attr + (size_t)g * 3 * D;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // the soft part adds to the same six slots from CTAs of the same launch: two addends on a zeroed slot, and
        // a + b == b + a, so the sum does not depend on who comes first
        atomicAdd(gpo + 2 * i + 0, inv * (-fk.q * A[i] + fk.p * Bv[i]));
        atomicAdd(gpo + 2 * i + 1, inv * (fk.n * A[i] - fk.m * Bv[i]));
    }
    if (P.attr_compact) {
        // nobody but dibr_backward_meshes reads dL/dattr, and it reads the depth channel alone: three floats per face
        if (dep) {
            float* __restrict__ g3 = P.grad_face_attr + (size_t)g * 3;
#pragma unroll
            for (int d = 0; d < DMAX; d++)
                if (d == nA + ones) { g3[0] = acc[0 * DMAX + d]; g3[1] = acc[1 * DMAX + d]; g3[2] = acc[2 * DMAX + d]; }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int d = 0; d < DMAX; d++)
            if (d < D) gao[i * D + d] = acc[i * DMAX + d];
}

// ---- soft part: one group of lanes per face that entered at least one soft-silhouette product ----------------------
#ifndef DIBR_SOFT_LANES
#define DIBR_SOFT_LANES 16
#endif
constexpr int SOFT_LANES = DIBR_SOFT_LANES;          // lanes per face: a face has ~12 contributing pixels out of ~80 scanned,
constexpr int SOFT_GROUPS = 256 / SOFT_LANES;        // so a full warp per face leaves most lanes idle in the evaluation
static_assert(SOFT_LANES == 8 || SOFT_LANES == 16 || SOFT_LANES == 32, "a sub-warp group");
__device__ __forceinline__ void backward_soft_body(const BwdParams& P, int bid)
{
    const int tid = threadIdx.x;
    const int grp = tid / SOFT_LANES, lane = tid % SOFT_LANES;          // lane = position inside the group
    const int shift = ((tid & 31) / SOFT_LANES) * SOFT_LANES;           // first lane of the group inside its warp
    const unsigned full = (SOFT_LANES == 32) ? 0xffffffffu : (((1u << SOFT_LANES) - 1u) << shift);      // the group's lanes
    const int wi = bid * SOFT_GROUPS + grp;
    if (wi >= P.list_counts[1]) return;                      // group-uniform
    const int W = P.width, H = P.height;
    const int g = P.soft_list[wi];
    const FaceRec rec = P.recs[g];
    const int b = __float_as_int(rec.image);
    const int f = __float_as_int(rec.local_id);
    const size_t img = (size_t)b * H * W;
    const int32_t* __restrict__ idx = P.imidx + img;
    const float* __restrict__ gpr = P.grad_improb + img;
    const float* __restrict__ comp = P.imcomp + img;
    // pixel centres inside the expanded bbox (record, written by the set-up kernel)
    const int c0 = (int)(rec.ecols & 0xffffu), c1 = (int)(rec.ecols >> 16);
    const int r0 = (int)(rec.erows & 0xffffu), r1 = (int)(rec.erows >> 16);
    const int nc = c1 - c0, npx = nc * (r1 - r0);
    if (nc <= 0 || npx <= 0) return;
    const float mult = (float)P.multiplier;
    const float zscale = (float)P.delta / (mult * mult);
    const float sentinel = 4.0f * mult * mult;
    const float px[3] = {rec.ax, rec.bx, rec.cx}, py[3] = {rec.ay, rec.by, rec.cy};
    float gp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto evaluate = [&](int r, int c) {
        const size_t pix = (size_t)r * W + c;
        const float x0 = P.xs[c], y0 = P.ys[r];
        const SoftHit h = soft_distance(rec.ax, rec.ay, rec.bx, rec.by, rec.cx, rec.cy, x0, y0, sentinel);
        float p, om;
        soft_prob(h.d2 * zscale, p, om);
        // d improb / d p_k = prod_{j != k}(1 - p_j) = comp / (1 - p_k);  dp/dz = -p;  z = zscale * d2;
        // the gradient w.r.t. UN-multiplied coordinates carries one more 'mult'
        const float coef = __fdividef(-__ldg(gpr + pix) * __ldg(comp + pix), om + 1e-15f) * p * zscale * mult;     // 2 ulp division: gradients are gated at 1e-5
        if (h.kase >= 3) {
            const int k = h.kase - 3;
            const float vx = (k == 0) ? px[0] : ((k == 1) ? px[1] : px[2]);
            const float vy = (k == 0) ? py[0] : ((k == 1) ? py[1] : py[2]);
            const float gx = coef * 2.0f * (vx - x0), gy = coef * 2.0f * (vy - y0);
            if (k == 0) { gp[0] += gx; gp[1] += gy; }
            else if (k == 1) { gp[2] += gx; gp[3] += gy; }
            else { gp[4] += gx; gp[5] += gy; }
        } else {
            const int k = h.kase, k2 = (k + 1) % 3;
            const float x1 = (k == 0) ? px[0] : ((k == 1) ? px[1] : px[2]);
            const float y1 = (k == 0) ? py[0] : ((k == 1) ? py[1] : py[2]);
            const float x2 = (k2 == 0) ? px[0] : ((k2 == 1) ? px[1] : px[2]);
            const float y2 = (k2 == 0) ? py[0] : ((k2 == 1) ? py[1] : py[2]);
            const float exx = x2 - x1, eyy = y2 - y1;
            const float s2 = __fdividef(2.0f * coef, h.len2);
            const float gx1 = s2 * ((y0 - y2) * h.cr + h.d2 * exx);
            const float gy1 = s2 * (-(x0 - x2) * h.cr + h.d2 * eyy);
            const float gx2 = s2 * (-(y0 - y1) * h.cr - h.d2 * exx);
            const float gy2 = s2 * ((x0 - x1) * h.cr - h.d2 * eyy);
            if (k == 0) { gp[0] += gx1; gp[1] += gy1; gp[2] += gx2; gp[3] += gy2; }
            else if (k == 1) { gp[2] += gx1; gp[3] += gy1; gp[4] += gx2; gp[5] += gy2; }
            else { gp[4] += gx1; gp[5] += gy1; gp[0] += gx2; gp[1] += gy2; }
        }
    };

    // The forward left one bit per pixel ("uncovered").  Lane = pixel row of the expanded range: it reads its row 32 pixels
    // per word and touches imidx only where the bit is set (a face has ~12 contributing pixels among the ~80 of its range).
    // The surviving (row, column) pairs of the group are then numbered by a prefix sum and dealt out evenly, one per lane
    // and turn.  Which lane adds which pixel is a function of the masks alone and the lanes are combined by a fixed tree:
    // bit-reproducible.
    const int wb = (W + 7) >> 3;
    const unsigned char* __restrict__ orow = P.open8 + (size_t)b * H * wb;
    const unsigned char* __restrict__ crow = P.closed8 + (size_t)b * H * wb;
    for (int rb = r0; rb < r1; rb += SOFT_LANES) {
        const int r = rb + lane;
        for (int cbase = c0 & ~7; cbase < c1; cbase += 32) {
            unsigned m = 0u;
            if (r < r1) {
                const size_t o8 = (size_t)r * wb + (cbase >> 3);
                unsigned mc = 0u;                               // uncovered pixels closed by their K-th face: imidx = -(that face + 1)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (cbase + 8 * k < c1 && (cbase >> 3) + k < wb) {
                        m |= (unsigned)__ldg(orow + o8 + k) << (8 * k);
                        mc |= (unsigned)__ldg(crow + o8 + k) << (8 * k);
                    }
                const int lo = max(c0 - cbase, 0), hi = min(c1 - cbase, 32);
                m &= ((hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                // an uncovered pixel that never reached K faces counted every near face; a closed one (about one in eight)
                // only the faces up to its K-th: imidx is read there alone
                for (unsigned t = m & mc; t; t &= t - 1) {
                    const int bit = __ffs(t) - 1;
                    const int v = idx[(size_t)r * W + cbase + bit];
                    if (!(v == 0 || f + 1 <= -v)) m &= ~(1u << bit);
                }
            }
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < SOFT_LANES; o <<= 1) {
                const int t = __shfl_up_sync(full, incl, o, SOFT_LANES);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(full, incl, SOFT_LANES - 1, SOFT_LANES);
            for (int jb = 0; jb < total; jb += SOFT_LANES) {
                const int jq = jb + lane;
                int src = 0;                                    // lanes whose running count is <= jq: the lane that holds pair jq
#pragma unroll
                for (int st = SOFT_LANES / 2; st >= 1; st >>= 1) {
                    const int v = __shfl_sync(full, incl, min(src + st - 1, SOFT_LANES - 1), SOFT_LANES);
                    if (v <= jq) src += st;
                }
                src = min(src, SOFT_LANES - 1);
                const unsigned ms = __shfl_sync(full, m, src, SOFT_LANES);
                const int before = __shfl_sync(full, incl - cnt, src, SOFT_LANES);
                if (jq < total) evaluate(rb + src, cbase + (int)__fns(ms, 0u, jq - before + 1));
            }
        }
    }
    // fixed-tree reduction over the group, then add to the face's six slots
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float v = gp[i];
#pragma unroll
        for (int o = SOFT_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
        gp[i] = v;
    }
    {   // every lane of the group holds the totals: lanes 0..5 deliver one each
        float v = gp[0];
#pragma unroll
        for (int i = 1; i < 6; i++) v = (lane == i) ? gp[i] : v;
        if (lane < 6) atomicAdd(P.grad_points2d + (size_t)g * 6 + lane, v);
    }
}

// One persistent launch for both parts.  The list lengths live on the device, so a grid sized for the worst case would
// be mostly CTAs that find nothing to do; instead a fixed grid strides over the work items that exist: first the
// colour items (32 faces each), then the soft items (256 / SOFT_LANES faces each) -- both kinds share the SMs as the first run out.
#ifndef DIBR_BWD_MIN_CTAS
#define DIBR_BWD_MIN_CTAS 4        // 64 registers: measured faster than 80 registers at 3 CTAs per SM
#endif
#ifndef DIBR_BWD_GRID
#define DIBR_BWD_GRID (148 * 16)
#endif
template <int DMAX>
__global__ void __launch_bounds__(256, (DMAX <= 8) ? DIBR_BWD_MIN_CTAS : 2) backward_faces_kernel(const __grid_constant__ BwdParams P, int do_color, int do_soft)
{
    // dL/dattr of the faces that won no pixel is zero.  The colour body writes the whole row of every face on its list, so
    // only the OTHER rows are cleared here (about half of them, inside this launch) instead of a memset of the whole array.
    if (!P.attr_compact) {           // (compact mode: prepare_backward_kernel has zeroed the [total_faces, 3] depth column)
        const unsigned int* __restrict__ flags = reinterpret_cast<const unsigned int*>(P.face_flags);
        const int row = 3 * P.num_attr;                                     // floats per face
        const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
        if ((row & 3) == 0 && (reinterpret_cast<uintptr_t>(P.grad_face_attr) & 15) == 0) {
            const int r4 = row >> 2;
            const long long n4 = (long long)P.total_faces * r4;
            for (long long i = t0; i < n4; i += stride) {
                const int g = (int)(i / r4);
                if (!do_color || !(__ldg(flags + g) & 1u)) reinterpret_cast<float4*>(P.grad_face_attr)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const long long n = (long long)P.total_faces * row;
            for (long long i = t0; i < n; i += stride) {
                const int g = (int)(i / row);
                if (!do_color || !(__ldg(flags + g) & 1u)) P.grad_face_attr[i] = 0.f;
            }
        }
    }
    const int cbn = do_color ? (P.list_counts[0] + (256 / GRP) - 1) / (256 / GRP) : 0;
    const int sbn = do_soft ? (P.list_counts[1] + SOFT_GROUPS - 1) / SOFT_GROUPS : 0;
    for (int it = blockIdx.x; it < cbn + sbn; it += gridDim.x) {
        __syncwarp();                        // lanes leave the bodies at different points
        if (it < cbn) backward_color_body<DMAX>(P, it);
        else backward_soft_body(P, it - cbn);
    }
}

// Ahead of the face kernel, in the launch that used to be the memset of dL/dpoints2d: (1) zero dL/dpoints2d (faces on neither
// work list keep a zero gradient, the two bodies add into it); (2) compact the flags the forward left (bit 0: won a pixel,
// bit 1: entered a soft product) into the colour and the soft work list -- one counter atomic per CTA and list.  A face that
// has been listed gets bit 2 / bit 3, so a second backward over the same forward appends nothing and finds the lists as they are.
constexpr int PREP_THREADS = 256, PREP_FACES = 4 * PREP_THREADS;       // four faces per thread: one 16 B load of their flags
__global__ void __launch_bounds__(PREP_THREADS) prepare_backward_kernel(const BwdParams P)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = gridDim.x * PREP_THREADS, t0 = blockIdx.x * PREP_THREADS + tid;
    {
        const int n = 6 * P.total_faces;                    // (total_faces < 2^31 / 6: the ABI caps batch x faces well below)
        if ((reinterpret_cast<uintptr_t>(P.grad_points2d) & 15) == 0) {
            const int n4 = n >> 2;
            float4* __restrict__ o = reinterpret_cast<float4*>(P.grad_points2d);
            for (int i = t0; i < n4; i += nthreads) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = (n4 << 2) + t0; i < n; i += nthreads) P.grad_points2d[i] = 0.f;
        } else {
            for (int i = t0; i < n; i += nthreads) P.grad_points2d[i] = 0.f;
        }
    }
    if (P.attr_compact && (P.va.flags & 2)) {               // the compact depth column of dL/dattr: rows of unlisted faces stay zero
        const int n = 3 * P.total_faces;
        for (int i = t0; i < n; i += nthreads) P.grad_face_attr[i] = 0.f;
    }
    __shared__ int wcnt[2][PREP_THREADS / 32];
    __shared__ int cta_base[2];
    unsigned int* __restrict__ flags = reinterpret_cast<unsigned int*>(const_cast<unsigned char*>(P.face_flags));
    const bool vec = (reinterpret_cast<uintptr_t>(flags) & 15) == 0;
    for (int g0 = blockIdx.x * PREP_FACES; g0 < P.total_faces; g0 += gridDim.x * PREP_FACES) {          // uniform across the CTA
        const int g = g0 + 4 * tid;
        unsigned f[4] = {0u, 0u, 0u, 0u};
        if (vec && g + 3 < P.total_faces) {
            const uint4 v = *reinterpret_cast<const uint4*>(flags + g);
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (g + k < P.total_faces) f[k] = flags[g + k];
        }
        int nc = 0, ns = 0;                                 // this thread's new entries of the colour / soft list
#pragma unroll
        for (int k = 0; k < 4; k++) { nc += ((f[k] & 5u) == 1u) ? 1 : 0; ns += ((f[k] & 10u) == 2u) ? 1 : 0; }
        int ic = nc, is = ns;                               // inclusive scans over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, ic, o), b2 = __shfl_up_sync(0xffffffffu, is, o);
            if (lane >= o) { ic += a; is += b2; }
        }
        if (lane == 31) { wcnt[0][warp] = ic; wcnt[1][warp] = is; }
        __syncthreads();
        if (tid < 2) {                                      // one counter atomic per CTA and list
            int total = 0;
#pragma unroll
            for (int w = 0; w < PREP_THREADS / 32; w++) total += wcnt[tid][w];
            cta_base[tid] = total > 0 ? atomicAdd(&P.list_counts[tid], total) : 0;
        }
        __syncthreads();
        int oc = cta_base[0] + ic - nc, os = cta_base[1] + is - ns;
        for (int w = 0; w < warp; w++) { oc += wcnt[0][w]; os += wcnt[1][w]; }
        if (nc | ns) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool c = (f[k] & 5u) == 1u, s2 = (f[k] & 10u) == 2u;
                if (c) P.color_list[oc++] = g + k;
                if (s2) P.soft_list[os++] = g + k;
                if (c || s2) flags[g + k] = f[k] | (c ? 4u : 0u) | (s2 ? 8u : 0u);
            }
        }
        __syncthreads();                                    // wcnt / cta_base are rewritten by the next round
    }
}

int launch_backward_faces(const BwdParams& P, cudaStream_t stream, int parts)
{
    if (P.total_faces <= 0) return 0;
    if (P.total_faces > (1 << 28)) return (int)cudaErrorInvalidValue;       // 6 * total_faces is indexed with int
    cudaError_t e = cudaSuccess;
    if (parts & 1) {
        prepare_backward_kernel<<<min((P.total_faces + PREP_FACES - 1) / PREP_FACES, 148 * 8), PREP_THREADS, 0, stream>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    if (!(parts & 2)) return 0;
    const int do_color = P.any_grad_im ? 1 : 0, do_soft = (P.grad_improb && P.knum > 0) ? 1 : 0;
    if (!do_color && !do_soft) {                     // no upstream gradient at all: everything is zero
        if (P.attr_compact) return 0;                // (the prepare kernel has zeroed the compact column)
        e = cudaMemsetAsync(P.grad_face_attr, 0, sizeof(float) * 3 * (size_t)P.num_attr * (size_t)P.total_faces, stream);
        return (int)e;
    }
    const int worst = (do_color ? (P.total_faces + (256 / GRP) - 1) / (256 / GRP) : 0) + (do_soft ? (P.total_faces + SOFT_GROUPS - 1) / SOFT_GROUPS : 0);
    const int grid = min(worst, DIBR_BWD_GRID);
    if (P.num_attr <= 4) backward_faces_kernel<4><<<grid, 256, 0, stream>>>(P, do_color, do_soft);
    else if (P.num_attr <= 8) backward_faces_kernel<8><<<grid, 256, 0, stream>>>(P, do_color, do_soft);
    else backward_faces_kernel<12><<<grid, 256, 0, stream>>>(P, do_color, do_soft);
    return (int)cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// Fused-mode tail: per-face gradients -> per-vertex (fixed-order gather over the vertex's incident
// (face,corner) list) -> through divide / projection / view transform -> dL/d cam_view_R, dL/d cam_view_pos
// (and optionally dL/d vertices, dL/d vertex attributes).  Replaces the torch autograd graph the
// reference keeps for index_select, cat, division and matmul (perpsective.py:80-101, vcrender_batch.py:84-88).
// -------------------------------------------------------------------------------------------------
// (128-thread CTAs, which make the 16 x num_instances grid resident in one wave, measured slower: 22.5 vs 20.5 us)
#ifndef DIBR_MV_THREADS
#define DIBR_MV_THREADS 256
#endif
constexpr int MV_T = DIBR_MV_THREADS;
__global__ void __launch_bounds__(MV_T) mesh_vertex_grad_kernel(MeshBwdParams P)
{
    const int inst = blockIdx.y;
    const int32_t* de = P.inst_desc + inst * INST_STRIDE;
    // an instance without faces has no camera in the workspace (the set-up kernel's first face writes it) and no gradient
    const int nv = (de[I_NUM_FACES] > 0) ? de[I_NUM_VERTS] : 0, vbase = de[I_VERT_BASE], fbase = de[I_OUT_FACE_BASE], gvbase = de[I_GVERT_BASE];
    const int abase = de[I_ADJ_BASE];      // row of this mesh's vertex 0 in the CSR pointer array
    const float* R = P.cam_rot + (size_t)de[I_CAM] * 9;
    const float* T = P.cam_pos + (size_t)de[I_CAM] * 3;
    const float* Pm = P.cam_proj + (size_t)de[I_PROJ] * 16;
    const int A = P.vert_attr_dim, D = P.num_attr;
    const int depth_ch = (P.attr_flags & 2) ? (A + ((P.attr_flags & 1) ? 1 : 0)) : -1;

    float s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0.f;

    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        const int e0 = P.vert_face_ptr[abase + v], e1 = P.vert_face_ptr[abase + v + 1];
        float g2x = 0.f, g2y = 0.f, gdep = 0.f;
        float ga[DIBR_MAX_ATTR_INTERNAL];
#pragma unroll
        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) ga[d] = 0.f;
        // incident (face, corner) pairs, 8 at a time: all indices first, then all gradient loads, then the sums in
        // list order (the order fixes the rounding; the loads overlap instead of chaining L2 round trips)
        for (int eb = e0; eb < e1; eb += 8) {
            int fcs[8];
#pragma unroll
            for (int k = 0; k < 8; k++) fcs[k] = (eb + k < e1) ? __ldg(P.vert_face_idx + eb + k) : -1;
            float2 g2[8];
            float gdk[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                g2[k] = make_float2(0.f, 0.f); gdk[k] = 0.f;
                if (fcs[k] >= 0) {
                    const int gf = fbase + fcs[k] / 3, c = fcs[k] % 3;
                    g2[k] = *reinterpret_cast<const float2*>(P.grad_points2d + (size_t)gf * 6 + c * 2);
                    if (depth_ch >= 0) gdk[k] = P.attr_compact ? P.grad_face_attr[(size_t)gf * 3 + c] : P.grad_face_attr[((size_t)gf * 3 + c) * D + depth_ch];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (fcs[k] >= 0) {
                    g2x += g2[k].x; g2y += g2[k].y; gdep += gdk[k];
                    if (P.grad_vert_attr) {
                        const float* gfa = P.grad_face_attr + ((size_t)(fbase + fcs[k] / 3) * 3 + fcs[k] % 3) * D;
#pragma unroll
                        for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) if (d < A) ga[d] += gfa[d];
                    }
                }
            }
        }
        if (P.grad_vert_attr) {
            float* o = P.grad_vert_attr + (size_t)(gvbase + v) * A;
#pragma unroll
            for (int d = 0; d < DIBR_MAX_ATTR_INTERNAL; d++) if (d < A) o[d] = ga[d];
        }
        // forward recomputation (cheap) for the chain rule
        const float* vp = P.verts + (size_t)(vbase + v) * P.verts_stride;
        const float d0 = vp[0] - T[0], d1 = vp[1] - T[1], d2 = vp[2] - T[2];
        float pc[3];
#pragma unroll
        for (int j = 0; j < 3; j++) pc[j] = fmaf(R[j * 3 + 2], d2, fmaf(R[j * 3 + 1], d1, R[j * 3 + 0] * d0));
        const float cxv = fmaf(pc[2], Pm[8 + 0], fmaf(pc[1], Pm[4 + 0], pc[0] * Pm[0])) + Pm[12 + 0];
        const float cyv = fmaf(pc[2], Pm[8 + 1], fmaf(pc[1], Pm[4 + 1], pc[0] * Pm[1])) + Pm[12 + 1];
        const float cwv = fmaf(pc[2], Pm[8 + 3], fmaf(pc[1], Pm[4 + 3], pc[0] * Pm[3])) + Pm[12 + 3];
        const float iw = 1.0f / cwv;
        const float gcx = g2x * iw, gcy = g2y * iw;
        const float gcw = -(g2x * cxv + g2y * cyv) * iw * iw;
        float gpc[3];
#pragma unroll
        for (int r = 0; r < 3; r++) gpc[r] = Pm[r * 4 + 0] * gcx + Pm[r * 4 + 1] * gcy + Pm[r * 4 + 3] * gcw;
        gpc[2] -= gdep;                                  // depth attribute = -z_view
        float gd[3];
#pragma unroll
        for (int k = 0; k < 3; k++) gd[k] = R[0 * 3 + k] * gpc[0] + R[1 * 3 + k] * gpc[1] + R[2 * 3 + k] * gpc[2];
        if (P.grad_verts) {
            float* o = P.grad_verts + (size_t)(gvbase + v) * 3;
            o[0] = gd[0]; o[1] = gd[1]; o[2] = gd[2];
        }
        const float dd[3] = {d0, d1, d2};
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) s[j * 3 + k] += gpc[j] * dd[k];
        s[9] -= gd[0]; s[10] -= gd[1]; s[11] -= gd[2];
    }
    // fixed-tree block reduction of the 12 pose sums
    __shared__ float red[MV_T / 32][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        float v = s[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < MV_T / 32; w++) v += red[w][threadIdx.x];
        P.pose_part[((size_t)inst * gridDim.x + blockIdx.x) * 12 + threadIdx.x] = v;
    }
    // ---- the block that delivers last sums the partials of its instance in block order (fixed, so the result does not
    //      depend on which block that is) and, in pose mode, chains cam_view_R = F R (F = diag(1,-1,-1)) and
    //      cam_view_pos = -(R^T t) down to R and t:
    //        dL/dR[j][k] = F_jj dL/dcamR[j][k] - t[j] dL/dpos[k],   dL/dt[j] = -sum_k R[j][k] dL/dpos[k]
    __shared__ unsigned int ticket;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();            // one fence for the block (cumulative over the barrier): the partial sums before the ticket
        ticket = atomicAdd(&P.pose_done[inst], 1u);
    }
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    if (threadIdx.x == 0) P.pose_done[inst] = 0u;          // ready for another backward over the same forward
    __threadfence();
    __shared__ float tot[12];
    if (threadIdx.x < 12) {
        const int nblocks = gridDim.x;
        float v = 0.f;
        for (int b0 = 0; b0 < nblocks; b0 += 8) {
            float t[8];
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = (b0 + k < nblocks) ? __ldcg(P.pose_part + ((size_t)inst * nblocks + b0 + k) * 12 + threadIdx.x) : 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) v += t[k];
        }
        tot[threadIdx.x] = v;
        if (threadIdx.x < 9) { if (P.grad_cam_rot) P.grad_cam_rot[(size_t)inst * 9 + threadIdx.x] = v; }
        else if (P.grad_cam_pos) P.grad_cam_pos[(size_t)inst * 3 + threadIdx.x - 9] = v;
    }
    __syncthreads();
    if (P.pose_R && threadIdx.x < 12) {
        const float* Rp = P.pose_R + (size_t)inst * 9;
        const float* Tp = P.pose_t + (size_t)inst * 3;
        if (threadIdx.x < 9) {
            const int j = threadIdx.x / 3, k = threadIdx.x - 3 * j;
            const float sgn = (j == 0) ? 1.f : -1.f;
            const float v = sgn * tot[j * 3 + k] - Tp[j] * tot[9 + k];
            P.grad_pose_R[(size_t)inst * 9 + threadIdx.x] = v;
            if (P.grad_pose_packed) P.grad_pose_packed[(size_t)inst * 12 + threadIdx.x] = v;
            if (P.host_pose_packed) P.host_pose_packed[(size_t)inst * 12 + threadIdx.x] = v;
        } else {
            const int j = threadIdx.x - 9;
            float gt = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) gt -= Rp[j * 3 + k] * tot[9 + k];
            P.grad_pose_t[(size_t)inst * 3 + j] = gt;
            if (P.grad_pose_packed) P.grad_pose_packed[(size_t)inst * 12 + 9 + j] = gt;
            if (P.host_pose_packed) P.host_pose_packed[(size_t)inst * 12 + 9 + j] = gt;
        }
    }
    // ---- optional row num_instances of the packed gradients: their column sums, added in instance order by the block that
    //      finalises the LAST instance -- the 12-float vector a data-parallel step all-reduces, ready without another launch
    if (P.pose_sum && P.grad_pose_packed && P.pose_R) {
        __shared__ unsigned int t2;
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); t2 = atomicAdd(&P.pose_done[P.num_instances], 1u); }
        __syncthreads();
        if (t2 != (unsigned)P.num_instances - 1u) return;
        if (threadIdx.x == 0) P.pose_done[P.num_instances] = 0u;
        __threadfence();
        if (threadIdx.x < 12) {
            float v = 0.f;
            for (int i = 0; i < P.num_instances; i++) v += __ldcg(P.grad_pose_packed + (size_t)i * 12 + threadIdx.x);
            P.grad_pose_packed[(size_t)P.num_instances * 12 + threadIdx.x] = v;
        }
    }
}

int launch_backward_meshes(const MeshBwdParams& P, cudaStream_t stream)
{
    if (P.num_instances <= 0) return 0;
    dim3 grid(POSE_BLOCKS, P.num_instances);
    mesh_vertex_grad_kernel<<<grid, MV_T, 0, stream>>>(P);        // the last block of every instance finalises it
    return (int)cudaGetLastError();
}

}  // namespace dibr
