// Post-render crop & resize: batch_crop_resize (core/utils/zoom_utils.py:80-95), i.e. ROIAlign(output_size, 1.0, 0,
// aligned=True) of detectron2.layers.roi_align (un-vendored; a thin wrapper over torchvision.ops.roi_align, whose
// published algorithm is restated here).  Self6D++ renders full frames and crops the colour image / the teacher normals
// to the network's ROI with it (self_engine_utils.py:528-533, 662-666, 690-692).
//
//   forward   one thread per output pixel (roi, ph, pw), channels in groups of four that reuse the sample weights.  The
//             input is addressed through element strides, so the renderer's channels-last images are read in place (the
//             reference permutes and the op copies to contiguous first).
//   backward  the reference scatters with fp32 atomicAdd (order dependent).  Here the input gradient is GATHERED: one
//             thread per input pixel walks, in ascending roi order, the sample rows and columns whose bilinear footprint
//             holds the pixel -- the sample grid of a roi is regular, so they form one index range per axis, estimated
//             with a margin and then evaluated with exactly the forward's expressions -- and adds
//             wy * wx * g / count in a fixed order: bit-reproducible, and the dense gradient is written once (no memset).
#include "dibr_common.cuh"
#include "dibr_internal.h"

namespace dibr {

namespace {

struct RoiGeom {
    int img;                    // batch index
    float start_w, start_h;     // roi corner (after scale and the half-pixel offset)
    float bin_w, bin_h;         // size of one output bin
    int grid_w, grid_h;         // samples per bin and axis
};

__device__ __forceinline__ RoiGeom roi_geom(const RoiAlignParams& P, int r)
{
    const float* q = P.rois + (size_t)r * 5;
    RoiGeom g;
    g.img = (int)q[0];
    const float off = P.aligned ? 0.5f : 0.0f;
    g.start_w = q[1] * P.spatial_scale - off;
    g.start_h = q[2] * P.spatial_scale - off;
    const float end_w = q[3] * P.spatial_scale - off, end_h = q[4] * P.spatial_scale - off;
    float rw = end_w - g.start_w, rh = end_h - g.start_h;
    if (!P.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    g.bin_h = rh / (float)P.pooled_h;
    g.bin_w = rw / (float)P.pooled_w;
    g.grid_h = P.sampling_ratio > 0 ? P.sampling_ratio : (int)ceilf(rh / (float)P.pooled_h);
    g.grid_w = P.sampling_ratio > 0 ? P.sampling_ratio : (int)ceilf(rw / (float)P.pooled_w);
    return g;
}

// one axis of bilinear_interpolate: sample coordinate -> (low, high, weight of low, weight of high); false = outside
__device__ __forceinline__ bool axis_taps(float v, int size, int& lo, int& hi, float& w_lo, float& w_hi)
{
    if (v < -1.0f || v > (float)size) return false;
    if (v <= 0.f) v = 0.f;
    lo = (int)v;
    if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
    const float l = v - (float)lo;
    w_lo = 1.0f - l; w_hi = l;
    return true;
}

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid)
{   // roi_start + ph * bin_size + (iy + .5f) * bin_size / roi_bin_grid -- no contraction, so the forward and the backward
    // (and the CPU restatement) see the same coordinate
    return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}

constexpr int RA_T = 256;
constexpr int RA_CG = 4;        // channels per pass

__global__ void __launch_bounds__(RA_T) roi_align_forward_kernel(RoiAlignParams P)
{
    const long long total = (long long)P.num_rois * P.pooled_h * P.pooled_w;
    for (long long i = (long long)blockIdx.x * RA_T + threadIdx.x; i < total; i += (long long)gridDim.x * RA_T) {
        const int pw = (int)(i % P.pooled_w);
        const long long t = i / P.pooled_w;
        const int ph = (int)(t % P.pooled_h);
        const int r = (int)(t / P.pooled_h);
        const RoiGeom g = roi_geom(P, r);
        const float count = (float)max(g.grid_h * g.grid_w, 1);
        const float* base = P.input + (long long)g.img * P.stride_n;
        float* out = P.output + ((size_t)r * P.channels * P.pooled_h + ph) * P.pooled_w + pw;
        const size_t out_cs = (size_t)P.pooled_h * P.pooled_w;
        const bool img_ok = g.img >= 0 && g.img < P.num_images;
        for (int c0 = 0; c0 < P.channels; c0 += RA_CG) {
            float acc[RA_CG];
#pragma unroll
            for (int k = 0; k < RA_CG; k++) acc[k] = 0.f;
            if (img_ok) {
                for (int iy = 0; iy < g.grid_h; iy++) {
                    int y0, y1; float hy, ly;
                    const bool oky = axis_taps(sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h), P.height, y0, y1, hy, ly);
                    for (int ix = 0; ix < g.grid_w; ix++) {
                        int x0, x1; float hx, lx;
                        const bool okx = axis_taps(sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w), P.width, x0, x1, hx, lx);
                        if (!(oky && okx)) continue;
                        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                        const long long o1 = y0 * P.stride_h + x0 * P.stride_w, o2 = y0 * P.stride_h + x1 * P.stride_w;
                        const long long o3 = y1 * P.stride_h + x0 * P.stride_w, o4 = y1 * P.stride_h + x1 * P.stride_w;
#pragma unroll
                        for (int k = 0; k < RA_CG; k++) {
                            if (c0 + k < P.channels) {
                                const float* pl = base + (long long)(c0 + k) * P.stride_c;
                                acc[k] += w1 * __ldg(pl + o1) + w2 * __ldg(pl + o2) + w3 * __ldg(pl + o3) + w4 * __ldg(pl + o4);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RA_CG; k++)
                if (c0 + k < P.channels) out[(size_t)(c0 + k) * out_cs] = acc[k] / count;
        }
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------
#ifndef DIBR_RA_TILE_H
#define DIBR_RA_TILE_H 32
#endif
constexpr int BT_W = 64, BT_H = DIBR_RA_TILE_H;   // input tile of one CTA
constexpr int BT_ROWS = RA_T / BT_W;        // rows the CTA covers at a time
constexpr int BT_PIX = BT_H / BT_ROWS;      // pixels per thread
constexpr int RA_MAXC = 8;                  // column taps of a pixel kept in registers (more: recomputed in the loop)
constexpr int LISTCAP = 512;                // rois of one tile kept in shared memory between flushes (>= RA_T)

// candidate sample indices j (over pooled * grid samples of one axis) whose footprint can hold pixel v: conservative
__device__ __forceinline__ void sample_range(float start, float bin, int grid, int pooled, int v, int size, int& j0, int& j1)
{
    const int n = pooled * grid;
    const float d = bin / (float)grid;                           // sample pitch
    // a pitch below 1e-3 px (a roi of less than a pixel), or nan: examine all samples -- the estimate below is good to
    // ~1e-4 / d samples (fp32 rounding of v - start), far inside its margin of one sample otherwise
    if (!(d > 1e-3f)) { j0 = 0; j1 = n - 1; return; }
    // unclamped samples reach v when they lie in (v-1, v+1); the borders also take the clamped ones in [-1,0] / [size-1,size]
    const float lo = (v == 0) ? -1.5f : (float)v - 1.0f, hi = (v == size - 1) ? (float)size + 0.5f : (float)v + 1.0f;
    const float inv = 1.0f / d;
    const float a = (lo - start) * inv - 0.5f, b = (hi - start) * inv - 0.5f;
    j0 = (a < -1.0f) ? 0 : ((a > (float)n) ? n : (int)a - 1);
    j1 = (b < -1.0f) ? -1 : ((b > (float)n) ? n - 1 : (int)b + 2);
    j0 = max(j0, 0); j1 = min(j1, n - 1);
}

#ifndef DIBR_RA_MIN_CTAS
#define DIBR_RA_MIN_CTAS 3          /* measured 2 / 3 / 4: 0.208 / 0.154 / 0.166 ms for the 256x256 crops of 32 frames */
#endif
__global__ void __launch_bounds__(RA_T, DIBR_RA_MIN_CTAS) roi_align_backward_kernel(RoiAlignParams P)
{
    __shared__ int s_list[LISTCAP];
    __shared__ int s_wcount[RA_T / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (P.width + BT_W - 1) / BT_W, tiles_y = (P.height + BT_H - 1) / BT_H;
    const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
    const int tx0 = (tile % tiles_x) * BT_W, ty0 = (tile / tiles_x) * BT_H;
    // BT_PIX pixels per thread, BT_ROWS rows apart: the per-tile roi search (a global load, divisions, two barriers) is
    // paid once per 2048 pixels -- with one pixel per thread the kernel was bound by that latency (0.26 ms for 118 MB)
    const int x = tx0 + (tid % BT_W), yb = ty0 + (tid / BT_W);
    const size_t g_cs = (size_t)P.pooled_h * P.pooled_w;

    for (int c0 = 0; c0 < P.channels; c0 += RA_CG) {
        float* gin = P.grad_input + (long long)n * P.stride_n + (long long)x * P.stride_w + (long long)c0 * P.stride_c;
        // adds the listed rois (ascending) to this thread's pixels; `first` starts from zero, later rois / flushes continue
        // from what the same thread stored (read back through L1/L2): one fixed order per pixel.  Per roi the thread's
        // column taps are found once for its BT_PIX pixels; the row taps are the same for the whole warp (a warp is half a
        // tile row), so lane t evaluates row candidate t and the loop broadcasts them.
        auto flush = [&](int nl, bool first) {
            const bool col_ok = x < P.width;
            if (first) {
                // ---- zeros for the whole tile first (most tiles of a frame are reached by no roi and end here)
                const bool linear = P.stride_c == 1 && P.stride_w == P.channels;      // channels-last: a tile row is one run
                if (linear) {
                    if (c0 == 0) {
                        const int run = min(BT_W, P.width - tx0) * P.channels;
                        float* tile0 = P.grad_input + (long long)n * P.stride_n + (long long)tx0 * P.stride_w;
                        const bool vec = ((run | (int)(P.stride_h & 3) | (int)(P.stride_n & 3)) & 3) == 0 &&
                                         (reinterpret_cast<uintptr_t>(P.grad_input) & 15) == 0;   // tx0 * channels is a multiple of 64
                        for (int i = 0; i < BT_PIX; i++) {
                            const int y = yb + i * BT_ROWS;
                            if (y >= P.height) break;
                            float* row = tile0 + (long long)y * P.stride_h;
                            if (vec) {
                                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                                for (int e = tid % BT_W; e < (run >> 2); e += BT_W) reinterpret_cast<float4*>(row)[e] = z;
                            } else {
                                for (int e = tid % BT_W; e < run; e += BT_W) row[e] = 0.f;
                            }
                        }
                    }
                    if (nl > 0) __syncthreads();                 // uniform; other threads read these zeros back below
                } else if (col_ok) {
                    for (int i = 0; i < BT_PIX; i++) {
                        const int y = yb + i * BT_ROWS;
                        if (y >= P.height) break;
#pragma unroll
                        for (int k = 0; k < RA_CG; k++)
                            if (c0 + k < P.channels) gin[(long long)y * P.stride_h + (long long)k * P.stride_c] = 0.f;
                    }
                }
            }
            for (int li = 0; li < nl; li++) {
                const int rr = s_list[li];
                const RoiGeom g = roi_geom(P, rr);
                const float inv_count = 1.0f / (float)max(g.grid_h * g.grid_w, 1);
                const float* go = P.grad_output + (size_t)rr * P.channels * g_cs + (size_t)c0 * g_cs;
                // ---- column taps of this thread: sample columns [jxs, jxs + ncx) with weights cwx (0 where a candidate misses)
                int jx0 = 0, jx1 = -1;
                if (col_ok) sample_range(g.start_w, g.bin_w, g.grid_w, P.pooled_w, x, P.width, jx0, jx1);
                float cwx[RA_MAXC];
                int jxs = 0, ncx = 0, nhit = 0;
                bool spill = false;                              // more than RA_MAXC columns: weights recomputed in the loop
                for (int jx = jx0; jx <= jx1; jx++) {
                    const int pw = jx / g.grid_w, ix = jx - pw * g.grid_w;
                    int b0, b1; float wb0, wb1;
                    float wx = 0.f;
                    bool hit = false;
                    if (axis_taps(sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w), P.width, b0, b1, wb0, wb1) && (b0 == x || b1 == x)) {
                        wx = (b0 == x ? wb0 : 0.f) + (b1 == x ? wb1 : 0.f);
                        hit = true;
                    }
                    if (ncx == 0) { if (!hit) continue; jxs = jx; }
                    if (ncx < RA_MAXC) {
#pragma unroll
                        for (int c = 0; c < RA_MAXC; c++) if (c == ncx) cwx[c] = wx;
                    } else if (hit) spill = true;
                    if (hit || ncx < RA_MAXC) ncx = min(ncx + 1, RA_MAXC + 1);
                    if (hit) nhit = ncx;
                }
                if (!spill) ncx = nhit;                          // drop the trailing candidates that miss
                const int jx_last = jx1;
                if (!__any_sync(0xffffffffu, ncx > 0)) continue;     // the roi misses this warp's columns
                for (int i = 0; i < BT_PIX; i++) {
                    const int y = yb + i * BT_ROWS;              // uniform across the warp
                    if (y >= P.height) break;
                    int jy0, jy1;
                    sample_range(g.start_h, g.bin_h, g.grid_h, P.pooled_h, y, P.height, jy0, jy1);
                    float* o = gin + (long long)y * P.stride_h;
                    float acc[RA_CG];
#pragma unroll
                    for (int k = 0; k < RA_CG; k++) acc[k] = 0.f;
                    bool any = false;                            // some tap of this roi holds the pixel
                    for (int jb = jy0; jb <= jy1; jb += 32) {
                        const int jy = jb + lane;
                        int my_ph = 0;
                        float my_wy = 0.f;
                        if (jy <= jy1) {
                            const int ph = jy / g.grid_h, iy = jy - ph * g.grid_h;
                            int a0, a1; float wa0, wa1;
                            if (axis_taps(sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h), P.height, a0, a1, wa0, wa1) && (a0 == y || a1 == y)) {
                                my_wy = (a0 == y ? wa0 : 0.f) + (a1 == y ? wa1 : 0.f);
                                my_ph = ph;
                            }
                        }
                        const int nin = min(32, jy1 - jb + 1);
                        for (int t = 0; t < nin; t++) {
                            const float wy = __shfl_sync(0xffffffffu, my_wy, t);
                            const int ph = __shfl_sync(0xffffffffu, my_ph, t);
                            if (wy == 0.f) continue;             // uniform
                            const float* grow = go + (size_t)ph * P.pooled_w;
                            any = any || ncx > 0;
                            if (!spill) {
#pragma unroll
                                for (int c = 0; c < RA_MAXC; c++) {
                                    if (c < ncx) {
                                        const int jx = jxs + c;
                                        const int pw = g.grid_w == 1 ? jx : jx / g.grid_w;
                                        const float w = wy * cwx[c] * inv_count;
#pragma unroll
                                        for (int k = 0; k < RA_CG; k++)
                                            if (c0 + k < P.channels) acc[k] = fmaf(w, __ldg(grow + (size_t)k * g_cs + pw), acc[k]);
                                    }
                                }
                            } else {
                                for (int jx = jxs; jx <= jx_last; jx++) {
                                    const int pw = jx / g.grid_w, ix = jx - pw * g.grid_w;
                                    int b0, b1; float wb0, wb1;
                                    if (!axis_taps(sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w), P.width, b0, b1, wb0, wb1)) continue;
                                    if (b0 != x && b1 != x) continue;
                                    const float w = wy * ((b0 == x ? wb0 : 0.f) + (b1 == x ? wb1 : 0.f)) * inv_count;
#pragma unroll
                                    for (int k = 0; k < RA_CG; k++)
                                        if (c0 + k < P.channels) acc[k] = fmaf(w, __ldg(grow + (size_t)k * g_cs + pw), acc[k]);
                                }
                            }
                        }
                    }
                    if (col_ok && any) {                         // earlier rois' sum (or the zero) + this roi's, same thread every time
#pragma unroll
                        for (int k = 0; k < RA_CG; k++)
                            if (c0 + k < P.channels) o[(long long)k * P.stride_c] += acc[k];
                    }
                }
            }
        };
        bool first = true;
        int r0 = 0;
        while (true) {
            int nl = 0;
            while (r0 < P.num_rois) {
                // ---- rois of this image whose sample extent reaches the tile, ascending (ordered compaction)
                const int r = r0 + tid;
                bool take = false;
                if (r < P.num_rois) {
                    const RoiGeom g = roi_geom(P, r);
                    if (g.img == n && g.grid_h > 0 && g.grid_w > 0) {
                        const float ex = g.start_w + g.bin_w * (float)P.pooled_w, ey = g.start_h + g.bin_h * (float)P.pooled_h;
                        const float xa = fminf(g.start_w, ex) - 2.f, xb = fmaxf(g.start_w, ex) + 2.f;
                        const float ya = fminf(g.start_h, ey) - 2.f, yb2 = fmaxf(g.start_h, ey) + 2.f;
                        // a sample reaches pixels at most one away (also when clamped at the border): 2 is a safe margin
                        take = xb >= (float)tx0 && xa <= (float)(tx0 + BT_W) && yb2 >= (float)ty0 && ya <= (float)(ty0 + BT_H);
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, take);
                if (lane == 0) s_wcount[warp] = __popc(bal);
                __syncthreads();
                int base = 0, cnt_round = 0;
#pragma unroll
                for (int w = 0; w < RA_T / 32; w++) { const int cnt = s_wcount[w]; if (w < warp) base += cnt; cnt_round += cnt; }
                const bool full = nl + cnt_round > LISTCAP;      // uniform; the round is examined again after the flush
                if (!full && take) s_list[nl + base + __popc(bal & ((1u << lane) - 1u))] = r;
                __syncthreads();                                 // list entries visible; s_wcount free for the next round
                if (full) break;
                nl += cnt_round;
                r0 += RA_T;
            }
            flush(nl, first);
            first = false;
            if (r0 >= P.num_rois) break;
            __syncthreads();                                     // everybody is done reading the list
        }
        __syncthreads();                                         // the list is rebuilt for the next channel group
    }
}

// ---- RoIPool: batch_crop_resize(interpolation="nearest") = torchvision.ops.RoIPool(output_size, 1.0)
//      (/root/reference/core/utils/zoom_utils.py:91-92).  Quantised roi (roundf of the scaled corners, width / height
//      >= 1), bin (ph, pw) = rows floor(ph * bin_h) .. ceil((ph + 1) * bin_h) (shifted by the roi start, clipped to the
//      image), maximum and its first position in row-major order; an empty bin gives 0 and no gradient.
struct PoolBox { int b, sw, sh; float bin_h, bin_w; };
__device__ __forceinline__ PoolBox pool_box(const RoiPoolParams& P, int r) {
    const float* roi = P.rois + (size_t)r * 5;
    PoolBox B;
    B.b = (int)roi[0];
    B.sw = (int)roundf(roi[1] * P.spatial_scale); B.sh = (int)roundf(roi[2] * P.spatial_scale);
    const int ew = (int)roundf(roi[3] * P.spatial_scale), eh = (int)roundf(roi[4] * P.spatial_scale);
    const int rw = max(ew - B.sw + 1, 1), rh = max(eh - B.sh + 1, 1);               // malformed rois become 1 x 1
    B.bin_h = (float)rh / (float)P.pooled_h; B.bin_w = (float)rw / (float)P.pooled_w;
    return B;
}
__device__ __forceinline__ void pool_bin(const RoiPoolParams& P, const PoolBox& B, int ph, int pw, int& h0, int& h1, int& w0, int& w1) {
    h0 = min(max((int)floorf((float)ph * B.bin_h) + B.sh, 0), P.height);
    h1 = min(max((int)ceilf((float)(ph + 1) * B.bin_h) + B.sh, 0), P.height);
    w0 = min(max((int)floorf((float)pw * B.bin_w) + B.sw, 0), P.width);
    w1 = min(max((int)ceilf((float)(pw + 1) * B.bin_w) + B.sw, 0), P.width);
}

__global__ void __launch_bounds__(256) roi_pool_forward_kernel(const RoiPoolParams P)
{
    const long long total = (long long)P.num_rois * P.channels * P.pooled_h * P.pooled_w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int pw = (int)(i % P.pooled_w), ph = (int)((i / P.pooled_w) % P.pooled_h);
        const int c = (int)((i / ((long long)P.pooled_w * P.pooled_h)) % P.channels);
        const int r = (int)(i / ((long long)P.pooled_w * P.pooled_h * P.channels));
        const PoolBox B = pool_box(P, r);
        int h0, h1, w0, w1;
        pool_bin(P, B, ph, pw, h0, h1, w0, w1);
        const bool empty = (h1 <= h0) || (w1 <= w0) || B.b < 0 || B.b >= P.num_images;
        float best = empty ? 0.f : -3.402823466e+38f;
        int arg = -1;
        if (!empty) {
            const float* img = P.input + (long long)B.b * P.stride_n + (long long)c * P.stride_c;
            for (int h = h0; h < h1; h++)
                for (int w = w0; w < w1; w++) {
                    const float v = img[(long long)h * P.stride_h + (long long)w * P.stride_w];
                    if (v > best) { best = v; arg = h * P.width + w; }
                }
        }
        P.output[i] = best;
        P.argmax[i] = arg;
    }
}

// Gradient by gather, one thread per input element: the bins of every roi of its image that can hold the pixel are
// visited in (roi, ph, pw) order and those whose recorded maximum IS the pixel add their gradient -- no atomics, every
// element written once, bit-reproducible (torchvision scatters with atomicAdd).
__global__ void __launch_bounds__(256) roi_pool_backward_kernel(const RoiPoolParams P)
{
    const long long total = (long long)P.num_images * P.channels * P.height * P.width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % P.width), h = (int)((i / P.width) % P.height);
        const int c = (int)((i / ((long long)P.width * P.height)) % P.channels);
        const int b = (int)(i / ((long long)P.width * P.height * P.channels));
        float g = 0.f;
        for (int r = 0; r < P.num_rois; r++) {
            const PoolBox B = pool_box(P, r);
            if (B.b != b) continue;
            // candidate bins: bin ph holds row h' = h - start iff floor(ph bin) <= h' < ceil((ph + 1) bin), which puts ph strictly
            // between (h' - 1) / bin - 1 and (h' + 1) / bin; the recorded argmax decides membership exactly
            const float hr = (float)(h - B.sh), wr = (float)(w - B.sw);
            const int p0 = max((int)floorf((hr - 1.f) / B.bin_h) - 1, 0), p1 = min((int)floorf((hr + 1.f) / B.bin_h) + 1, P.pooled_h - 1);
            const int q0 = max((int)floorf((wr - 1.f) / B.bin_w) - 1, 0), q1 = min((int)floorf((wr + 1.f) / B.bin_w) + 1, P.pooled_w - 1);
            for (int ph = p0; ph <= p1; ph++)
                for (int pw = q0; pw <= q1; pw++) {
                    const size_t o = (((size_t)r * P.channels + c) * P.pooled_h + ph) * P.pooled_w + pw;
                    if (P.argmax[o] == h * P.width + w) g += P.grad_output[o];
                }
        }
        P.grad_input[(long long)b * P.stride_n + (long long)c * P.stride_c + (long long)h * P.stride_h + (long long)w * P.stride_w] = g;
    }
}

}  // namespace

int launch_roi_pool_forward(const RoiPoolParams& P, cudaStream_t stream)
{
    const long long total = (long long)P.num_rois * P.channels * P.pooled_h * P.pooled_w;
    if (total <= 0) return 0;
    const long long want = (total + 255) / 256;
    roi_pool_forward_kernel<<<(int)(want < 148ll * 32 ? want : 148ll * 32), 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

int launch_roi_pool_backward(const RoiPoolParams& P, cudaStream_t stream)
{
    const long long total = (long long)P.num_images * P.channels * P.height * P.width;
    if (total <= 0) return 0;
    const long long want = (total + 255) / 256;
    roi_pool_backward_kernel<<<(int)(want < 148ll * 32 ? want : 148ll * 32), 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

int launch_roi_align_forward(const RoiAlignParams& P, cudaStream_t stream)
{
    const long long total = (long long)P.num_rois * P.pooled_h * P.pooled_w;
    if (total <= 0 || P.channels <= 0) return 0;
    const long long want = (total + RA_T - 1) / RA_T;
    const int grid = (int)(want < 148ll * 32 ? want : 148ll * 32);
    roi_align_forward_kernel<<<grid, RA_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

int launch_roi_align_backward(const RoiAlignParams& P, cudaStream_t stream)
{
    if (P.num_images <= 0 || P.channels <= 0 || P.height <= 0 || P.width <= 0) return 0;
    const long long tiles = (long long)((P.width + BT_W - 1) / BT_W) * ((P.height + BT_H - 1) / BT_H) * P.num_images;
    roi_align_backward_kernel<<<(unsigned)tiles, RA_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
