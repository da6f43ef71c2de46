// Helpers shared by the single-launch decode kernels (trm_decode.cu, bah_decode.cu): a cluster-wide split GEMV
// that leaves its result in every CTA's shared memory, and deterministic block reductions.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace ac {

namespace cg = cooperative_groups;
constexpr int kThreads = 1024;     // 32 warps: enough loads in flight to stream the weights from L2
constexpr int kWarps = kThreads / 32;

// out[r][n] = act(bias[n] + sum_k xin[r][k] * Wt[k][n]); xin/out/part in shared memory.
// Split-K inside the CTA: thread = (K-slice s, column quad c) streams its slice of 4 adjacent
// columns with independent 128-bit loads (a warp reads 512 contiguous bytes per k), partial sums go
// through `part` [KS][R][N] and are reduced in a fixed order (deterministic).  N % 4 == 0.
template <int R>
__device__ __forceinline__ void matvec_t(const float* __restrict__ Wt, const float* __restrict__ bias,
                                         const float* xin, int ldx, float* out, int ldo, int N, int K, bool relu,
                                         float* part, int rank, int P) {
    // this CTA owns column quads [rank * NCl, (rank + 1) * NCl) and writes them into EVERY CTA's `out`
    const int NCl = (N >> 2) / P, NC = N >> 2;
    const int Nl = NCl * 4;
    const int KS = max(1, min(kThreads / NCl, 4096 / Nl));
    const int kslice = (K + KS - 1) / KS;
    const int tid = threadIdx.x;
    const int s = tid / NCl, cl = tid - s * NCl;
    if (s < KS) {
        const int k0 = s * kslice, k1 = min(K, k0 + kslice);
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const float4* w = reinterpret_cast<const float4*>(Wt) + rank * NCl + cl;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) {
            const float4 wv = __ldg(w + (size_t)k * NC);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float xv = xin[r * ldx + k];
                acc[r][0] = fmaf(wv.x, xv, acc[r][0]); acc[r][1] = fmaf(wv.y, xv, acc[r][1]);
                acc[r][2] = fmaf(wv.z, xv, acc[r][2]); acc[r][3] = fmaf(wv.w, xv, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            *reinterpret_cast<float4*>(part + ((size_t)s * R + r) * Nl + 4 * cl) =
                make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = tid; i < R * Nl; i += kThreads) {
        const int r = i / Nl, nl = i - r * Nl;
        const int n = rank * Nl + nl;
        float v = bias ? __ldg(bias + n) : 0.0f;
        for (int q = 0; q < KS; ++q) v += part[((size_t)q * R + r) * Nl + nl];
        v = relu ? fmaxf(v, 0.0f) : v;
        for (int pr = 0; pr < P; ++pr) cluster.map_shared_rank(out, pr)[r * ldo + n] = v;   // own copy included
    }
}

__device__ __forceinline__ void block_argmax(float& v, int& idx, float* s_v, int* s_i) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { s_v[warp] = v; s_i[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        v = lane < kWarps ? s_v[lane] : -INFINITY;
        idx = lane < kWarps ? s_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) { s_v[0] = v; s_i[0] = idx; }
    }
    __syncthreads();
    v = s_v[0]; idx = s_i[0];
    __syncthreads();
}

__device__ __forceinline__ float block_sum(float v, float* s_v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) s_v[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += s_v[i];   // fixed order: deterministic
    __syncthreads();
    return t;
}


}  // namespace ac
