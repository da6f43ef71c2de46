// fp32 SIMT GEMM (see gemm.cuh).  Exact-parity path: plain fp32 FMAs, fp32 accumulate.
#include "gemm.cuh"

namespace ac {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_tn_kernel(GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_VECS = BM * BK / 4, B_VECS = BN * BK / 4;     // float4 loads per tile
    constexpr int A_PER = (A_VECS + NT - 1) / NT, B_PER = (B_VECS + NT - 1) / NT;
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M = g.M, N = g.N, K = g.K;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    float4 ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int v = tid + i * NT;
            int r = v / (BK / 4), kv = (v % (BK / 4)) * 4;
            int row = m0 + r, k = k0 + kv;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < A_VECS && row < M && k < K) {
                x = __ldg(reinterpret_cast<const float4*>(g.A + (size_t)row * (g.lda ? g.lda : K) + k));
                if (g.ascale != nullptr) {
                    float4 s = __ldg(reinterpret_cast<const float4*>(
                        g.ascale + (size_t)(row / g.rows_per_group) * K + k));
                    x.x *= s.x; x.y *= s.y; x.z *= s.z; x.w *= s.w;
                }
            }
            ra[i] = x;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int v = tid + i * NT;
            int r = v / (BK / 4), kv = (v % (BK / 4)) * 4;
            int col = n0 + r, k = k0 + kv;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < B_VECS && col < N && k < K)
                x = __ldg(reinterpret_cast<const float4*>(g.W + (size_t)col * K + k));
            rb[i] = x;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int v = tid + i * NT;
            if (v < A_VECS) {
                int r = v / (BK / 4), kv = (v % (BK / 4)) * 4;
                As[kv + 0][r] = ra[i].x; As[kv + 1][r] = ra[i].y;
                As[kv + 2][r] = ra[i].z; As[kv + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int v = tid + i * NT;
            if (v < B_VECS) {
                int r = v / (BK / 4), kv = (v % (BK / 4)) * 4;
                Bs[kv + 0][r] = rb[i].x; Bs[kv + 1][r] = rb[i].y;
                Bs[kv + 2][r] = rb[i].z; Bs[kv + 3][r] = rb[i].w;
            }
        }
    };

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4*>(a + i) = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4*>(b + j) = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int ldc = g.ldc ? g.ldc : N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int row = m0 + ty * TM + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            int col = n0 + tx * TN + j;
            if (col >= N) continue;
            float4 s = g.cscale ? __ldg(reinterpret_cast<const float4*>(g.cscale + col))
                                : make_float4(1.f, 1.f, 1.f, 1.f);
            float4 bb = g.cbias ? __ldg(reinterpret_cast<const float4*>(g.cbias + col))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            float v[4] = {fmaf(acc[i][j], s.x, bb.x), fmaf(acc[i][j + 1], s.y, bb.y),
                          fmaf(acc[i][j + 2], s.z, bb.z), fmaf(acc[i][j + 3], s.w, bb.w)};
            if (g.act == ACT_SWISH) {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = swishf(v[q]);
            } else if (g.act == ACT_RELU) {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.0f);
            }
            if (g.R != nullptr) {
                float4 r = __ldg(reinterpret_cast<const float4*>(g.R + (size_t)row * ldc + col));
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
            *reinterpret_cast<float4*>(g.C + (size_t)row * ldc + col) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

template <int BM, int BN, int TM, int TN>
static int launch(const GemmArgs& g, cudaStream_t st) {
    dim3 grid(cdiv(g.M, BM), cdiv(g.N, BN));
    static const std::string name = "gemm_tn_" + std::to_string(BM) + "x" + std::to_string(BN);
    AC_TIMED(name.c_str(), st);
    gemm_tn_kernel<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g);
    AC_LAUNCHED("gemm_tn_kernel");
    return AC_OK;
}

int gemm_tn(const GemmArgs& g, cudaStream_t st) {
    if (g.tw != nullptr && g.K % 8 == 0) return gemm_tc(g, st);
    return gemm_tn_simt(g, st);
}

int gemm_tn_simt(const GemmArgs& g, cudaStream_t st) {
    AC_REQUIRE(g.K % 4 == 0 && g.N % 4 == 0, "gemm_tn: K (%d) and N (%d) must be multiples of 4", g.K, g.N);
    AC_REQUIRE(g.M >= 0 && cdiv(g.N, 32) <= 65535, "gemm_tn: M/N out of range");
    if (g.M == 0 || g.N == 0) return AC_OK;
    // pick the largest tile that still gives >= 2 waves of CTAs on 148 SMs
    auto ctas = [&](int bm, int bn) { return (int64_t)cdiv(g.M, bm) * cdiv(g.N, bn); };
    const int64_t want = 2 * kNumSMs;
    if (g.N <= 32) {
        if (ctas(256, 32) >= want) return launch<256, 32, 8, 4>(g, st);
        return launch<64, 32, 4, 4>(g, st);   // 128 threads
    }
    if (g.N > 64 && ctas(128, 128) >= want) return launch<128, 128, 8, 8>(g, st);
    if (ctas(128, 64) >= want) return launch<128, 64, 8, 4>(g, st);
    return launch<64, 64, 4, 4>(g, st);
}

}  // namespace ac
