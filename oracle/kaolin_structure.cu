/*
 * BENCH / TEST INFRASTRUCTURE ONLY -- not product code; nothing under self6dpp_b200/ links or calls this.
 *
 * GPU stand-in with the STRUCTURE of the kernels the reference calls through
 * kaolin.graphics.dib_renderer.cuda.rasterizer.forward / backward (kaolin v0.1, un-vendored; call sites
 * /root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 and :249-269): one thread per PIXEL,
 * every thread loops over ALL faces of its image, the backward scatters with fp32 atomicAdd.  It is a direct CUDA
 * transcription of the algorithm restated in oracle/dibr_oracle_body.h (8(d)
 * "reference on the same box"), compiled for sm_100a, so that bench.py can put "the reference's own algorithm on
 * this GPU" beside the B200-native path (extra key kaolin_structure_gpu).  It is NOT kaolin's source.
 *
 * Also here: an FP32 FMA micro-benchmark (the FP32-side roofline denominator SURVEY.md 8(d) asks for).
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ float pix_x(int w, int width, int multiplier) { return (float)(1.0 * multiplier / width * (2 * w + 1 - width)); }
__device__ __forceinline__ float pix_y(int h, int height, int multiplier) { return (float)(1.0 * multiplier / height * (height - 2 * h - 1)); }

// K1: dr_cuda_forward_render_batch
__global__ void ks_forward_render(const float* __restrict__ points3d, const float* __restrict__ points2d, const float* __restrict__ direct,
                                  const float* __restrict__ bbox, const float* __restrict__ feat, float* imidx, float* imdep, float* imwei,
                                  float* im, int bnum, int height, int width, int fnum, int dnum, int multiplier)
{
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)bnum * height * width) return;
    const int wid = (int)(pix % width), hei = (int)((pix / width) % height), b = (int)(pix / ((long)width * height));
    const float x0 = pix_x(wid, width, multiplier), y0 = pix_y(hei, height, multiplier);
    float znow = imdep[pix];
    for (int f = 0; f < fnum; f++) {
        const long s1 = (long)b * fnum + f;
        if (direct[s1] < 0) continue;
        const float xmin = bbox[s1 * 4], ymin = bbox[s1 * 4 + 1], xmax = bbox[s1 * 4 + 2], ymax = bbox[s1 * 4 + 3];
        if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;
        const float ax = points2d[s1 * 6], ay = points2d[s1 * 6 + 1], bx = points2d[s1 * 6 + 2], by = points2d[s1 * 6 + 3];
        const float cx = points2d[s1 * 6 + 4], cy = points2d[s1 * 6 + 5];
        const float m = bx - ax, p = by - ay, n = cx - ax, q = cy - ay, s = x0 - ax, t = y0 - ay;
        const float k1 = s * q - n * t, k2 = m * t - s * p, k3 = m * q - n * p;
        const float w1 = (float)((double)k1 / ((double)k3 + 1e-15)), w2 = (float)((double)k2 / ((double)k3 + 1e-15));
        const float w0 = 1.0f - w1 - w2;
        if (w0 < 0 || w1 < 0 || w2 < 0) continue;
        const float z0 = w0 * points3d[s1 * 9 + 2] + w1 * points3d[s1 * 9 + 5] + w2 * points3d[s1 * 9 + 8];
        if (z0 <= znow) continue;
        znow = z0;
        imidx[pix] = f + 1.0f;
        imdep[pix] = z0;
        imwei[pix * 3] = w0; imwei[pix * 3 + 1] = w1; imwei[pix * 3 + 2] = w2;
        for (int d = 0; d < dnum; d++)
            im[pix * dnum + d] = w0 * feat[s1 * 3 * dnum + d] + w1 * feat[s1 * 3 * dnum + dnum + d] + w2 * feat[s1 * 3 * dnum + 2 * dnum + d];
    }
}

__device__ float face_min_dis(const float* p6, float x0, float y0, int multiplier, int* edgeid)
{
    float pdis[6];
    for (int i = 0; i < 3; i++) {
        const float x1 = p6[i * 2], y1 = p6[i * 2 + 1], x2 = p6[((i + 1) % 3) * 2], y2 = p6[((i + 1) % 3) * 2 + 1];
        const float A = y2 - y1, B = x1 - x2, C = x2 * y1 - x1 * y2;
        const float up = A * x0 + B * y0 + C, down = A * A + B * B;
        float x3 = B * B * x0 - A * B * y0 - A * C, y3 = A * A * y0 - A * B * x0 - B * C;
        x3 = (float)((double)x3 / ((double)down + 1e-15));
        y3 = (float)((double)y3 / ((double)down + 1e-15));
        const float dirv = (x3 - x1) * (x3 - x2) + (y3 - y1) * (y3 - y2);
        pdis[i] = (dirv > 0) ? (float)(4 * multiplier * multiplier) : (float)((double)(up * up) / ((double)down + 1e-15));
    }
    for (int i = 0; i < 3; i++) {
        const float x1 = p6[i * 2], y1 = p6[i * 2 + 1];
        pdis[i + 3] = (x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1);
    }
    int eid = 0;
    float best = pdis[0];
    for (int i = 1; i < 6; i++) if (best > pdis[i]) { best = pdis[i]; eid = i; }
    *edgeid = eid;
    return best;
}

// K2: dr_cuda_forward_prob_batch
__global__ void ks_forward_prob(const float* __restrict__ points2d, const float* __restrict__ bbox2, const float* __restrict__ imidx,
                                float* probface, float* probcase, float* probdis, float* improb,
                                int bnum, int height, int width, int fnum, int knum, int multiplier, int sigmainv)
{
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)bnum * height * width) return;
    const int wid = (int)(pix % width), hei = (int)((pix / width) % height), b = (int)(pix / ((long)width * height));
    if ((int)(imidx[pix] + 0.5f) - 1 >= 0) { improb[pix] = 1.0f; return; }
    const float x0 = pix_x(wid, width, multiplier), y0 = pix_y(hei, height, multiplier);
    int kid = 0;
    for (int f = 0; f < fnum && kid < knum; f++) {
        const long s1 = (long)b * fnum + f;
        const float xmin = bbox2[s1 * 4], ymin = bbox2[s1 * 4 + 1], xmax = bbox2[s1 * 4 + 2], ymax = bbox2[s1 * 4 + 3];
        if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;
        int eid;
        const float d2 = face_min_dis(points2d + s1 * 6, x0, y0, multiplier, &eid);
        const float z = (float)sigmainv * d2 / (float)multiplier / (float)multiplier;
        probface[pix * knum + kid] = f + 1.0f;
        probcase[pix * knum + kid] = eid + 1.0f;
        probdis[pix * knum + kid] = expf(-z);
        kid++;
    }
    float allprob = 1.0f;
    for (int i = 0; i < kid; i++) allprob *= (1.0f - probdis[pix * knum + i]);
    improb[pix] = 1.0f - allprob;
}

// K3: dr_cuda_backward_color_batch (fp32 atomics, order undefined)
__global__ void ks_backward_color(const float* __restrict__ grad_im, const float* __restrict__ imidx, const float* __restrict__ imwei,
                                  const float* __restrict__ points2d, const float* __restrict__ feat, float* grad_points2d, float* grad_feat,
                                  int bnum, int height, int width, int fnum, int dnum, int multiplier)
{
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)bnum * height * width) return;
    const int wid = (int)(pix % width), hei = (int)((pix / width) % height), b = (int)(pix / ((long)width * height));
    const int fi = (int)(imidx[pix] + 0.5f) - 1;
    if (fi < 0) return;
    const float x0 = pix_x(wid, width, multiplier), y0 = pix_y(hei, height, multiplier);
    const long s1 = (long)b * fnum + fi;
    for (int i = 0; i < 3; i++)
        for (int d = 0; d < dnum; d++) atomicAdd(grad_feat + s1 * 3 * dnum + i * dnum + d, grad_im[pix * dnum + d] * imwei[pix * 3 + i]);
    const float ax = points2d[s1 * 6], ay = points2d[s1 * 6 + 1], bx = points2d[s1 * 6 + 2], by = points2d[s1 * 6 + 3];
    const float cx = points2d[s1 * 6 + 4], cy = points2d[s1 * 6 + 5];
    const float m = bx - ax, p = by - ay, n = cx - ax, q = cy - ay, s = x0 - ax, t = y0 - ay;
    const float k1 = s * q - n * t, k2 = m * t - s * p, k3 = m * q - n * p;
    const float dw1dm = -q * k1, dw1dn = -t * k3 + p * k1, dw1dp = n * k1, dw1dq = s * k3 - m * k1, dw1ds = q * k3, dw1dt = -n * k3;
    const float dw2dm = t * k3 - q * k2, dw2dn = p * k2, dw2dp = -s * k3 + n * k2, dw2dq = -m * k2, dw2ds = -p * k3, dw2dt = m * k3;
    const float dw1[6] = {-(dw1dm + dw1dn + dw1ds), -(dw1dp + dw1dq + dw1dt), dw1dm, dw1dp, dw1dn, dw1dq};
    const float dw2[6] = {-(dw2dm + dw2dn + dw2ds), -(dw2dp + dw2dq + dw2dt), dw2dm, dw2dp, dw2dn, dw2dq};
    for (int d = 0; d < dnum; d++) {
        const float c0 = feat[s1 * 3 * dnum + d], c1 = feat[s1 * 3 * dnum + dnum + d], c2 = feat[s1 * 3 * dnum + 2 * dnum + d];
        const float dldI = (float)((double)((float)multiplier * grad_im[pix * dnum + d]) / ((double)(k3 * k3) + 1e-15));
        for (int j = 0; j < 6; j++) atomicAdd(grad_points2d + s1 * 6 + j, dldI * ((c1 - c0) * dw1[j] + (c2 - c0) * dw2[j]));
    }
}

// K4: dr_cuda_backward_prob_batch
__global__ void ks_backward_prob(const float* __restrict__ grad_improb, const float* __restrict__ improb, const float* __restrict__ imidx,
                                 const float* __restrict__ probface, const float* __restrict__ probcase, const float* __restrict__ probdis,
                                 const float* __restrict__ points2d, float* grad_points2dprob,
                                 int bnum, int height, int width, int fnum, int knum, int multiplier, int sigmainv)
{
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)bnum * height * width) return;
    const int wid = (int)(pix % width), hei = (int)((pix / width) % height), b = (int)(pix / ((long)width * height));
    if ((int)(imidx[pix] + 0.5f) - 1 >= 0) return;
    const float x0 = pix_x(wid, width, multiplier), y0 = pix_y(hei, height, multiplier);
    const float dLdp = grad_improb[pix], allprob = improb[pix];
    for (int kid = 0; kid < knum; kid++) {
        const int fi = (int)(probface[pix * knum + kid] + 0.5f) - 1;
        if (fi < 0) break;
        const long s6 = ((long)b * fnum + fi) * 6;
        const float prob = probdis[pix * knum + kid];
        const float dLdz = (float)(-1.0 * sigmainv * dLdp * (1.0 - allprob) / (1.0 - prob + 1e-15) * prob);
        const int eid = (int)(probcase[pix * knum + kid] + 0.5f) - 1;
        if (eid >= 3) {
            const long ps = s6 + (eid - 3) * 2;
            atomicAdd(grad_points2dprob + ps, dLdz * 2 * (points2d[ps] - x0) / multiplier);
            atomicAdd(grad_points2dprob + ps + 1, dLdz * 2 * (points2d[ps + 1] - y0) / multiplier);
        } else {
            const long ps = s6 + eid * 2, ps2 = s6 + ((eid + 1) % 3) * 2;
            const float x1 = points2d[ps], y1 = points2d[ps + 1], x2 = points2d[ps2], y2 = points2d[ps2 + 1];
            const float A = y2 - y1, B = x1 - x2, C = x2 * y1 - x1 * y2;
            const float up = A * x0 + B * y0 + C, down = A * A + B * B;
            const float dis = (float)((double)(up * up) / ((double)down + 1e-15));
            const float dzdA = (float)((double)(2 * (x0 * up - dis * A)) / ((double)down + 1e-15));
            const float dzdB = (float)((double)(2 * (y0 * up - dis * B)) / ((double)down + 1e-15));
            const float dzdC = (float)((double)(2 * up) / ((double)down + 1e-15));
            atomicAdd(grad_points2dprob + ps, dLdz * (dzdB - y2 * dzdC) / multiplier);
            atomicAdd(grad_points2dprob + ps + 1, dLdz * (x2 * dzdC - dzdA) / multiplier);
            atomicAdd(grad_points2dprob + ps2, dLdz * (y1 * dzdC - dzdB) / multiplier);
            atomicAdd(grad_points2dprob + ps2 + 1, dLdz * (dzdA - x1 * dzdC) / multiplier);
        }
    }
}

// FP32 FMA throughput: 8 independent chains per thread
__global__ void fma_peak_kernel(float* out, int iters)
{
    float a[8];
    const float x = 1.0000001f, y = 1e-7f * (float)threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = (float)k + y;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 123.456f) out[0] = s;              // never true: keeps the loop alive
}

}  // namespace

extern "C" {

/* forward of LinearRasterizer (rasterizer.py:73-220) on the GPU with the reference kernels' structure; all pointers device */
int ks_forward(const float* points3d, const float* points2d_m, const float* direct, const float* bbox, const float* bbox2, const float* feat,
               float* imidx, float* imdep, float* imwei, float* probface, float* probcase, float* probdis, float* im, float* improb,
               int bnum, int height, int width, int fnum, int dnum, int knum, int multiplier, int sigmainv, void* stream)
{
    const long npix = (long)bnum * height * width;
    const int grid = (int)((npix + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    ks_forward_render<<<grid, 256, 0, st>>>(points3d, points2d_m, direct, bbox, feat, imidx, imdep, imwei, im, bnum, height, width, fnum, dnum, multiplier);
    ks_forward_prob<<<grid, 256, 0, st>>>(points2d_m, bbox2, imidx, probface, probcase, probdis, improb, bnum, height, width, fnum, knum, multiplier, sigmainv);
    return (int)cudaGetLastError();
}

/* backward of LinearRasterizer (rasterizer.py:222-291) */
int ks_backward(const float* grad_im, const float* grad_improb, const float* improb, const float* imidx, const float* imwei,
                const float* probface, const float* probcase, const float* probdis, const float* points2d_m, const float* feat,
                float* grad_points2d, float* grad_feat, float* grad_points2dprob,
                int bnum, int height, int width, int fnum, int dnum, int knum, int multiplier, int sigmainv, void* stream)
{
    const long npix = (long)bnum * height * width;
    const int grid = (int)((npix + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    ks_backward_color<<<grid, 256, 0, st>>>(grad_im, imidx, imwei, points2d_m, feat, grad_points2d, grad_feat, bnum, height, width, fnum, dnum, multiplier);
    ks_backward_prob<<<grid, 256, 0, st>>>(grad_improb, improb, imidx, probface, probcase, probdis, points2d_m, grad_points2dprob,
                                           bnum, height, width, fnum, knum, multiplier, sigmainv);
    return (int)cudaGetLastError();
}

/* FP32 FMA peak of the current device in TFLOP/s (2 flops per FMA), CUDA events, best of `reps` */
double ks_fma_peak_tflops(int reps)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 16, block = 512;
    double best = 0.0;
    for (int r = 0; r < reps + 2; r++) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<grid, block>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return best;
}

}  // extern "C"
