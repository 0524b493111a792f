// Depth map -> compacted point cloud, the step in front of the chamfer op on Self6D++'s depth loss
// (core/self6dpp/losses/depth_bp_chamfer_loss.py:27-36: backproject_th (lib/pysixd/misc.py:350-367), then boolean-mask
// indexing `pc[pc[:, :, 2] > 0]` per sample).  The reference does this per sample with a host sync for the mask; here the
// whole batch is one ordered compaction on the device:
//   count   valid pixels (depth > 0) per chunk of 1024 pixels
//   write   every chunk sums the counts of the chunks before it (<= H*W/1024 loads), compacts its own pixels in order
//           (warp ballots + one CTA scan) and writes X = (u - cx) d / fx, Y = (v - cy) d / fy, Z = d in the same fp32
//           operation order as the torch expression, plus the pixel -> row map the backward needs
// Rows keep the reference's order (row-major over the image), so everything downstream is index-compatible.
#include "dibr_internal.h"

namespace dibr {

constexpr int BP_T = 1024;

__global__ void __launch_bounds__(BP_T) bp_count_kernel(BackprojectParams P)
{
    const int b = blockIdx.y, npix = P.height * P.width;
    const int i = blockIdx.x * BP_T + threadIdx.x;
    const bool valid = (i < npix) && (P.depth[(size_t)b * npix + i] > 0.0f);
    const int n = __syncthreads_count(valid ? 1 : 0);
    if (threadIdx.x == 0) P.chunk_count[(size_t)b * gridDim.x + blockIdx.x] = n;
}

__global__ void __launch_bounds__(BP_T) bp_write_kernel(BackprojectParams P)
{
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int b = blockIdx.y, npix = P.height * P.width;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i = blockIdx.x * BP_T + t;
    // rows taken by the chunks before this one
    int before = 0;
    for (int c = t; c < (int)blockIdx.x; c += BP_T) before += P.chunk_count[(size_t)b * gridDim.x + c];
    before = __reduce_add_sync(0xffffffffu, before);
    if (lane == 0) wsum[warp] = before;
    __syncthreads();
    if (t == 0) { int s = 0; for (int w = 0; w < 32; w++) s += wsum[w]; base_s = s; }
    __syncthreads();
    const int base = base_s;
    __syncthreads();
    const float d = (i < npix) ? P.depth[(size_t)b * npix + i] : 0.0f;
    const bool valid = d > 0.0f;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; w++) off += wsum[w];
    int total = 0;
    for (int w = 0; w < 32; w++) total += wsum[w];
    if (i < npix) {
        int row = -1;
        if (valid) {
            row = off + __popc(bal & ((1u << lane) - 1u));
            const float* K = P.K + (size_t)(P.num_K > 1 ? b : 0) * 9;
            const int v = i / P.width, u = i - v * P.width;
            float* o = P.points + ((size_t)b * npix + row) * 3;
            o[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, K[2]), d), K[0]);
            o[1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, K[5]), d), K[4]);
            o[2] = d;
        }
        P.slot[(size_t)b * npix + i] = row;
    }
    if (blockIdx.x == gridDim.x - 1 && t == 0) P.count[b] = base + total;
}

// d loss / d depth of a valid pixel = gX (u - cx) / fx + gY (v - cy) / fy + gZ; 0 elsewhere
__global__ void __launch_bounds__(256) bp_backward_kernel(BackprojectParams P)
{
    const int b = blockIdx.y, npix = P.height * P.width;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npix) return;
    const int row = P.slot[(size_t)b * npix + i];
    float g = 0.0f;
    if (row >= 0) {
        const float* K = P.K + (size_t)(P.num_K > 1 ? b : 0) * 9;
        const int v = i / P.width, u = i - v * P.width;
        const float* gp = P.grad_points + ((size_t)b * npix + row) * 3;
        g = gp[0] * ((float)u - K[2]) / K[0] + gp[1] * ((float)v - K[5]) / K[4] + gp[2];
    }
    P.grad_depth[(size_t)b * npix + i] = g;
}

int launch_backproject(const BackprojectParams& P, cudaStream_t stream)
{
    const int npix = P.height * P.width;
    if (P.batch <= 0 || npix <= 0) return 0;
    const dim3 grid((npix + BP_T - 1) / BP_T, P.batch);
    bp_count_kernel<<<grid, BP_T, 0, stream>>>(P);
    bp_write_kernel<<<grid, BP_T, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

int launch_backproject_backward(const BackprojectParams& P, cudaStream_t stream)
{
    const int npix = P.height * P.width;
    if (P.batch <= 0 || npix <= 0) return 0;
    bp_backward_kernel<<<dim3((npix + 255) / 256, P.batch), 256, 0, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace dibr
