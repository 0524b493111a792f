// Chamfer nearest-neighbour distance: B200 replacement for the reference's
// core/csrc/torch_nndistance (nnd_cuda_kernel.cu:8-130 NmDistanceKernel, :164-183 NmDistanceGradKernel; CPU twin
// nnd_cpu.cpp:3-26, :59-133), which Self6D++ feeds with back-projected rendered / observed depth
// (core/self6dpp/losses/depth_bp_chamfer_loss.py:12-62).
//
// Same exhaustive search and the same answers -- squared distance ((dx*dx + dy*dy) + dz*dz) in fp32 without
// contraction, first minimum in ascending target order -- but (1) ONE launch covers a whole ragged batch (the
// reference loops over samples in Python and its grid then uses 16 blocks), (2) the targets are staged through shared
// memory as float4 and every thread keeps two queries in flight, (3) the backward is a deterministic gather (each point
// scans the other cloud's index array in ascending order) instead of fp32 atomicAdd scatter.
#include "dibr_internal.h"

namespace dibr {

constexpr int NND_THREADS = 256;
constexpr int NND_TILE = 1024;          // targets per shared-memory tile
constexpr int NND_Q = 2;                // queries per thread

__device__ __forceinline__ float sqdist_ref(float qx, float qy, float qz, float4 t) {
    const float x2 = __fsub_rn(t.x, qx), y2 = __fsub_rn(t.y, qy), z2 = __fsub_rn(t.z, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(x2, x2), __fmul_rn(y2, y2)), __fmul_rn(z2, z2));
}

// queries: cloud A of sample b = rows [offA[b], offA[b] + cntA[b]) of xyzA; targets likewise in cloud B
__global__ void __launch_bounds__(NND_THREADS) nnd_forward_kernel(NndParams P, int dir)
{
    __shared__ float4 tile[NND_TILE];
    const int b = blockIdx.y;
    const float* __restrict__ A = dir == 0 ? P.xyz1 : P.xyz2;
    const float* __restrict__ Bm = dir == 0 ? P.xyz2 : P.xyz1;
    const int offA = dir == 0 ? P.stride1 * b : P.stride2 * b, offB = dir == 0 ? P.stride2 * b : P.stride1 * b;
    const int* cA = dir == 0 ? P.count1 : P.count2;
    const int* cB = dir == 0 ? P.count2 : P.count1;
    const int n = cA ? min(cA[b], dir == 0 ? P.stride1 : P.stride2) : (dir == 0 ? P.stride1 : P.stride2);
    const int m = cB ? min(cB[b], dir == 0 ? P.stride2 : P.stride1) : (dir == 0 ? P.stride2 : P.stride1);
    float* __restrict__ dist = (dir == 0 ? P.dist1 : P.dist2) + offA;
    int* __restrict__ idx = (dir == 0 ? P.idx1 : P.idx2) + offA;
    const int q0 = (blockIdx.x * NND_THREADS + threadIdx.x) * NND_Q;
    if (blockIdx.x * NND_THREADS * NND_Q >= n) return;             // whole CTA beyond the cloud
    float qx[NND_Q], qy[NND_Q], qz[NND_Q], best[NND_Q];
    int bi[NND_Q];
#pragma unroll
    for (int u = 0; u < NND_Q; u++) {
        const int j = min(q0 + u, n - 1);
        qx[u] = A[(size_t)(offA + j) * 3 + 0]; qy[u] = A[(size_t)(offA + j) * 3 + 1]; qz[u] = A[(size_t)(offA + j) * 3 + 2];
        best[u] = 0.f; bi[u] = 0;
    }
    for (int k0 = 0; k0 < m; k0 += NND_TILE) {
        const int nt = min(NND_TILE, m - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < nt; k += NND_THREADS) {
            const float* t = Bm + (size_t)(offB + k0 + k) * 3;
            tile[k] = make_float4(t[0], t[1], t[2], 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < nt; k++) {
            const float4 t = tile[k];
#pragma unroll
            for (int u = 0; u < NND_Q; u++) {
                const float d = sqdist_ref(qx[u], qy[u], qz[u], t);
                if ((k0 + k) == 0 || d < best[u]) { best[u] = d; bi[u] = k0 + k; }      // nnd_cpu.cpp:17
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NND_Q; u++)
        if (q0 + u < n) { dist[q0 + u] = best[u]; idx[q0 + u] = bi[u]; }
}

// gradient of cloud A's points: own term 2 g_A[j] (a_j - b_idxA[j]) minus the pulls from every point k of cloud B
// whose nearest neighbour is j: 2 g_B[k] (b_k - a_j)   (nnd_cpu.cpp:94-130), gathered in ascending k
__global__ void __launch_bounds__(NND_THREADS) nnd_backward_kernel(NndParams P, int dir)
{
    __shared__ int tidx[NND_TILE];
    __shared__ float4 tpt[NND_TILE];          // xyz of the target point, w = its upstream gradient
    const int b = blockIdx.y;
    const float* __restrict__ A = dir == 0 ? P.xyz1 : P.xyz2;
    const float* __restrict__ Bm = dir == 0 ? P.xyz2 : P.xyz1;
    const int sA = dir == 0 ? P.stride1 : P.stride2, sB = dir == 0 ? P.stride2 : P.stride1;
    const int offA = sA * b, offB = sB * b;
    const int* cA = dir == 0 ? P.count1 : P.count2;
    const int* cB = dir == 0 ? P.count2 : P.count1;
    const int n = cA ? min(cA[b], sA) : sA, m = cB ? min(cB[b], sB) : sB;
    const float* __restrict__ gA = (dir == 0 ? P.graddist1 : P.graddist2) + offA;
    const float* __restrict__ gB = (dir == 0 ? P.graddist2 : P.graddist1) + offB;
    const int* __restrict__ idxA = (dir == 0 ? P.idx1 : P.idx2) + offA;
    const int* __restrict__ idxB = (dir == 0 ? P.idx2 : P.idx1) + offB;
    float* __restrict__ out = (dir == 0 ? P.gradxyz1 : P.gradxyz2) + (size_t)offA * 3;
    const int j = blockIdx.x * NND_THREADS + threadIdx.x;
    if (blockIdx.x * NND_THREADS >= sA) return;
    const bool live = j < n;
    float ax = 0.f, ay = 0.f, az = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    if (live) {
        ax = A[(size_t)(offA + j) * 3 + 0]; ay = A[(size_t)(offA + j) * 3 + 1]; az = A[(size_t)(offA + j) * 3 + 2];
        if (m > 0) {
            const int j2 = idxA[j];
            const float g = gA[j] * 2.0f;
            gx = g * (ax - Bm[(size_t)(offB + j2) * 3 + 0]);
            gy = g * (ay - Bm[(size_t)(offB + j2) * 3 + 1]);
            gz = g * (az - Bm[(size_t)(offB + j2) * 3 + 2]);
        }
    }
    for (int k0 = 0; k0 < m; k0 += NND_TILE) {
        const int nt = min(NND_TILE, m - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < nt; k += NND_THREADS) {
            const float* t = Bm + (size_t)(offB + k0 + k) * 3;
            tidx[k] = idxB[k0 + k];
            tpt[k] = make_float4(t[0], t[1], t[2], gB[k0 + k]);
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int k = 0; k < nt; k++) {
                if (tidx[k] == j) {
                    const float4 t = tpt[k];
                    const float g = t.w * 2.0f;          // explicit roundings: dibr_nnd_grid.cu must reproduce this sum
                    gx = __fsub_rn(gx, __fmul_rn(g, __fsub_rn(t.x, ax)));
                    gy = __fsub_rn(gy, __fmul_rn(g, __fsub_rn(t.y, ay)));
                    gz = __fsub_rn(gz, __fmul_rn(g, __fsub_rn(t.z, az)));
                }
            }
        }
    }
    if (j < sA) {       // padded rows get zeros
        out[(size_t)j * 3 + 0] = live ? gx : 0.f; out[(size_t)j * 3 + 1] = live ? gy : 0.f; out[(size_t)j * 3 + 2] = live ? gz : 0.f;
    }
}

int launch_nnd_forward(const NndParams& P, cudaStream_t stream)
{
    if (P.batch <= 0) return 0;
    if (P.stride1 > 0) {
        dim3 g1((P.stride1 + NND_THREADS * NND_Q - 1) / (NND_THREADS * NND_Q), P.batch);
        nnd_forward_kernel<<<g1, NND_THREADS, 0, stream>>>(P, 0);
    }
    if (P.stride2 > 0) {
        dim3 g2((P.stride2 + NND_THREADS * NND_Q - 1) / (NND_THREADS * NND_Q), P.batch);
        nnd_forward_kernel<<<g2, NND_THREADS, 0, stream>>>(P, 1);
    }
    return (int)cudaGetLastError();
}

int launch_nnd_backward(const NndParams& P, cudaStream_t stream)
{
    if (P.batch <= 0) return 0;
    if (P.stride1 > 0) {
        dim3 g1((P.stride1 + NND_THREADS - 1) / NND_THREADS, P.batch);
        nnd_backward_kernel<<<g1, NND_THREADS, 0, stream>>>(P, 0);
    }
    if (P.stride2 > 0) {
        dim3 g2((P.stride2 + NND_THREADS - 1) / NND_THREADS, P.batch);
        nnd_backward_kernel<<<g2, NND_THREADS, 0, stream>>>(P, 1);
    }
    return (int)cudaGetLastError();
}

}  // namespace dibr
