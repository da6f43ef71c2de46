// Training-step building blocks (see train_ops.cuh) + the loss and optimizer entry points of the C ABI:
//   ac_ls_ce_fwd_bwd   captioning/losses/loss.py:51-74 `LabelSmoothingLoss.forward` and its gradient, fused
//   ac_clip_adam       python_scripts/train_eval/run.py:125-127 `clip_grad_norm_` + `torch.optim.Adam.step`
// All fp32.  The GEMMs run on the tcgen05 kernel of gemm_tc.cu (3xTF32); everything else here is HBM/latency-bound
// row-wise work over [rows, 256]-sized tensors (a training batch is 32 captions x <= 21 tokens = 672 rows).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "train_ops.cuh"

namespace ac {

// ------------------------------------------------------------------------------------ column sums / transpose
// 32 columns x 32 row lanes per block, 4 independent loads in flight per thread (the first version -- 8 row lanes, one load
// at a time -- took 14 us per launch, 41 launches per training step); fixed summation order: deterministic.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ X, int M, int N, int ld, float* __restrict__ out) {
    __shared__ float s[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + x;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    if (n < N) {
        int m = y;
        for (; m + 96 < M; m += 128) {
            a0 += X[(size_t)m * ld + n]; a1 += X[(size_t)(m + 32) * ld + n];
            a2 += X[(size_t)(m + 64) * ld + n]; a3 += X[(size_t)(m + 96) * ld + n];
        }
        for (; m < M; m += 32) a0 += X[(size_t)m * ld + n];
    }
    s[y][x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (y == 0 && n < N) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += s[i][x];
        out[n] = t;
    }
}
int colsum(const float* X, int M, int N, int ld, float* out, cudaStream_t st) {
    if (N <= 0) return AC_OK;
    colsum_kernel<<<cdiv(N, 32), 1024, 0, st>>>(X, M, N, ld, out);
    AC_LAUNCHED("colsum_kernel");
    return AC_OK;
}

__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ XT, int N, int Mp, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const float4* row = reinterpret_cast<const float4*>(XT + (size_t)n * Mp);
    float a0 = 0.0f, a1 = 0.0f;
    int i = lane;
    for (; i + 32 < Mp / 4; i += 64) {
        const float4 u = row[i], v = row[i + 32];
        a0 += (u.x + u.y) + (u.z + u.w); a1 += (v.x + v.y) + (v.z + v.w);
    }
    if (i < Mp / 4) { const float4 u = row[i]; a0 += (u.x + u.y) + (u.z + u.w); }
    float t = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[n] = t;
}
int rowsum(const float* XT, int N, int Mp, float* out, cudaStream_t st) {
    if (N <= 0) return AC_OK;
    AC_REQUIRE(Mp % 4 == 0, "rowsum: row length %d is not a multiple of 4", Mp);
    rowsum_kernel<<<cdiv(N, 8), 256, 0, st>>>(XT, N, Mp, out);
    AC_LAUNCHED("rowsum_kernel");
    return AC_OK;
}

// ------------------------------------------------------------------------------------ side streams (weight gradients)
int SideStreams::init() {
    const char* e = getenv("AC_TRAIN_SIDE");
    if (e != nullptr && atoi(e) == 0) return AC_OK;
    for (int i = 0; i < kN; ++i) {
        AC_CUDA(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
        AC_CUDA(cudaEventCreateWithFlags(&fork_ev[i], cudaEventDisableTiming));
        AC_CUDA(cudaEventCreateWithFlags(&done_ev[i], cudaEventDisableTiming));
    }
    AC_CUDA(cudaEventCreateWithFlags(&mark_ev, cudaEventDisableTiming));
    ready = true;
    return AC_OK;
}
void SideStreams::destroy() {
    for (int i = 0; i < kN; ++i) {
        if (s[i]) { cudaStreamSynchronize(s[i]); cudaStreamDestroy(s[i]); s[i] = nullptr; }
        if (fork_ev[i]) { cudaEventDestroy(fork_ev[i]); fork_ev[i] = nullptr; }
        if (done_ev[i]) { cudaEventDestroy(done_ev[i]); done_ev[i] = nullptr; }
    }
    if (mark_ev) { cudaEventDestroy(mark_ev); mark_ev = nullptr; }
    ready = false;
}
int SideStreams::fork(cudaStream_t main_st, cudaStream_t* side) {
    const int i = next; next = (next + 1) % kN;
    AC_CUDA(cudaEventRecord(fork_ev[i], main_st));
    AC_CUDA(cudaStreamWaitEvent(s[i], fork_ev[i], 0));
    pending[i] = true; last = i;
    *side = s[i];
    return AC_OK;
}
int SideStreams::mark() {
    AC_REQUIRE(last >= 0, "SideStreams::mark before any fork");
    AC_CUDA(cudaEventRecord(mark_ev, s[last]));
    return AC_OK;
}
int SideStreams::wait_mark(cudaStream_t main_st) {
    AC_CUDA(cudaStreamWaitEvent(main_st, mark_ev, 0));
    return AC_OK;
}
int SideStreams::join(cudaStream_t main_st) {
    for (int i = 0; i < kN; ++i) {
        if (!pending[i]) continue;
        AC_CUDA(cudaEventRecord(done_ev[i], s[i]));
        AC_CUDA(cudaStreamWaitEvent(main_st, done_ev[i], 0));
        pending[i] = false;
    }
    return AC_OK;
}

__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ X, int M, int N, int ld,
                                                            float* __restrict__ XT, int Mp) {
    __shared__ float tile[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = y; i < 32; i += 8) {
        const int m = m0 + i, n = n0 + x;
        tile[i][x] = (m < M && n < N) ? X[(size_t)m * ld + n] : 0.0f;
    }
    __syncthreads();
    for (int i = y; i < 32; i += 8) {
        const int n = n0 + i, m = m0 + x;
        if (n < N && m < Mp) XT[(size_t)n * Mp + m] = tile[x][i];
    }
}
int transpose_pad(const float* X, int M, int N, int ld, float* XT, int Mp, cudaStream_t st) {
    transpose_pad_kernel<<<dim3(cdiv(Mp, 32), cdiv(N, 32)), 256, 0, st>>>(X, M, N, ld, XT, Mp);
    AC_LAUNCHED("transpose_pad_kernel");
    return AC_OK;
}

// ------------------------------------------------------------------------------------ Linear
static inline int pad8(int m) { return (m + 7) / 8 * 8; }

size_t linear_pack_floats(int N, int K, bool need_dx) {
    return align_up(tc_packed_floats(N, K), 32) + (need_dx ? align_up(tc_packed_floats(K, N), 32) : 0);
}
int linear_refresh(Linear& l, bool need_dx, cudaStream_t st) {
    int rc = tc_pack_weight(l.W, nullptr, l.N, l.K, l.pk, st, &l.tw);
    if (rc != AC_OK) return rc;
    if (need_dx) {
        l.pkT = l.pk + align_up(tc_packed_floats(l.N, l.K), 32);
        rc = tc_pack_weight_strided(l.W, nullptr, l.K, l.N, 1, l.K, l.pkT, st, &l.twT);   // (k, n) -> W[n * K + k]
    }
    return rc;
}
// plan the (re-)pack of this layer's images as jobs of one batched launch; returns the next `first`
long long linear_plan(Linear& l, bool need_dx, long long first, std::vector<TcPackJob>& jobs) {
    TcPackJob j;
    first += tc_pack_plan(l.W, l.N, l.K, l.K, 1, l.pk, first, &j, &l.tw);
    jobs.push_back(j);
    if (need_dx) {
        l.pkT = l.pk + align_up(tc_packed_floats(l.N, l.K), 32);
        first += tc_pack_plan(l.W, l.K, l.N, 1, l.K, l.pkT, first, &j, &l.twT);     // (k, n) -> W[n * K + k]
        jobs.push_back(j);
    }
    return first;
}
int linear_fwd(const Linear& l, const float* X, int M, float* Y, int ldy, int act, const float* R, cudaStream_t st) {
    GemmArgs g;
    g.A = X; g.W = l.W; g.C = Y; g.M = M; g.N = l.N; g.K = l.K; g.cbias = l.b; g.act = act; g.R = R; g.ldc = ldy; g.tw = &l.tw;
    return gemm_tc(g, st);
}
size_t linear_bwd_scratch_floats(int M, int N, int K) {
    const int Mp = pad8(M);
    return align_up((size_t)N * Mp, 32) + align_up(tc_packed_floats(K, Mp), 32);
}
int linear_bwd(const Linear& l, const float* X, int ldx, const float* dY, int ldy, int M, float* dX, const float* R,
               float* scratch, cudaStream_t st, SideStreams* side) {
    int rc = AC_OK;
    const int Mp = pad8(M);
    if (l.dW != nullptr) {
        float* dYT = scratch;                                        // [N, Mp]
        float* xpk = scratch + align_up((size_t)l.N * Mp, 32);       // packed X^T: "weight" [K, Mp]
        rc = transpose_pad(dY, M, l.N, ldy, dYT, Mp, st); if (rc) return rc;
        cudaStream_t ws = st;                                        // the stream of the weight-gradient work
        if (side != nullptr && side->enabled()) { rc = side->fork(st, &ws); if (rc) return rc; }
        if (l.db != nullptr) { rc = rowsum(dYT, l.N, Mp, l.db, ws); if (rc) return rc; }
        TcWeight txw;
        rc = tc_pack_weight_strided(X, nullptr, l.K, M, 1, ldx, xpk, ws, &txw); if (rc) return rc;   // (k, m) -> X[m * ldx + k]
        txw.K = Mp;                                                  // columns M..Mp-1 of the pack are zero
        GemmArgs g;
        g.A = dYT; g.W = nullptr; g.C = l.dW; g.M = l.N; g.N = l.K; g.K = Mp; g.tw = &txw;
        rc = gemm_tc(g, ws); if (rc) return rc;
    } else if (l.db != nullptr) {
        rc = colsum(dY, M, l.N, ldy, l.db, st); if (rc) return rc;
    }
    if (dX != nullptr) {
        AC_REQUIRE(l.pkT != nullptr, "linear_bwd: the transposed weight was not packed (need_dx)");
        GemmArgs g;
        g.A = dY; g.W = nullptr; g.C = dX; g.M = M; g.N = l.K; g.K = l.N; g.R = R; g.tw = &l.twT; g.lda = ldy;
        rc = gemm_tc(g, st); if (rc) return rc;
    }
    return AC_OK;
}

// ------------------------------------------------------------------------------------ embedding
__global__ void embed_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ pe, const int64_t* __restrict__ word,
                                 int row0, int L, int D, int V, float scale, Dropout dp, float* __restrict__ X0) {
    const int m = row0 + blockIdx.x;
    const int64_t w = min((int64_t)V - 1, max((int64_t)0, word[m]));
    const int t = m % L;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const uint64_t idx = (uint64_t)m * D + d;
        float x = emb[(size_t)w * D + d] * drop_scale(dp.seed, 0, idx, dp.p) * scale + pe[(size_t)t * D + d];
        X0[idx] = x * drop_scale(dp.seed, 1, idx, dp.p);
    }
}
int embed_fwd(const float* emb, const float* pe, const int64_t* word, int row0, int n_rows, int L, int D, int V, float scale,
              Dropout dp, float* X0, cudaStream_t st) {
    if (n_rows <= 0) return AC_OK;
    embed_fwd_kernel<<<n_rows, 256, 0, st>>>(emb, pe, word, row0, L, D, V, scale, dp, X0);
    AC_LAUNCHED("embed_fwd_kernel");
    return AC_OK;
}
__global__ void embed_bwd_kernel(const float* __restrict__ dX0, const int64_t* __restrict__ word, int D, int V, float scale,
                                 Dropout dp, float* __restrict__ demb) {
    const int m = blockIdx.x;
    const int64_t w = min((int64_t)V - 1, max((int64_t)0, word[m]));
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const uint64_t idx = (uint64_t)m * D + d;
        const float g = dX0[idx] * drop_scale(dp.seed, 1, idx, dp.p) * scale * drop_scale(dp.seed, 0, idx, dp.p);
        if (g != 0.0f) atomicAdd(demb + (size_t)w * D + d, g);
    }
}
int embed_bwd(const float* dX0, const int64_t* word, int n_rows, int L, int D, int V, float scale, Dropout dp, float* demb,
              cudaStream_t st) {
    (void)L;
    if (n_rows <= 0) return AC_OK;
    embed_bwd_kernel<<<n_rows, 256, 0, st>>>(dX0, word, D, V, scale, dp, demb);
    AC_LAUNCHED("embed_bwd_kernel");
    return AC_OK;
}

// ------------------------------------------------------------------------------------ residual + LayerNorm
constexpr int kLnMaxPerLane = 8;   // D <= 256
__global__ void __launch_bounds__(128) add_ln_fwd_kernel(const float* __restrict__ X, const float* __restrict__ O,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         int row0, int M, int D, Dropout dp, uint32_t site,
                                                         float* __restrict__ S, float* __restrict__ mean,
                                                         float* __restrict__ rstd, float* __restrict__ Y) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + warp;
    if (r >= M) return;
    const int m = row0 + r;
    const int per = D / 32;
    float v[kLnMaxPerLane];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) {
        if (k < per) {
            const uint64_t idx = (uint64_t)m * D + lane + 32 * k;
            float s = O[idx] * drop_scale(dp.seed, site, idx, dp.p);
            if (X != nullptr) s += X[idx];
            v[k] = s; sum += s;
        }
    }
    sum = warp_sum(sum);
    const float mu = sum / D;
    float var = 0.0f;
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k)
        if (k < per) { const float d = v[k] - mu; var += d * d; }
    var = warp_sum(var) / D;
    const float rs = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) {
        if (k < per) {
            const int d = lane + 32 * k;
            const uint64_t idx = (uint64_t)m * D + d;
            if (S != nullptr) S[idx] = v[k];
            Y[idx] = (v[k] - mu) * rs * gamma[d] + beta[d];
        }
    }
    if (lane == 0) {
        if (mean != nullptr) mean[m] = mu;
        if (rstd != nullptr) rstd[m] = rs;
    }
}
int add_ln_fwd(const float* X, const float* O, const float* gamma, const float* beta, int row0, int M, int D, Dropout dp,
               uint32_t site, float* S, float* mean, float* rstd, float* Y, cudaStream_t st) {
    AC_REQUIRE(D % 32 == 0 && D <= 32 * kLnMaxPerLane, "add_ln_fwd: width %d not supported", D);
    if (M <= 0) return AC_OK;
    add_ln_fwd_kernel<<<cdiv(M, 4), 128, 0, st>>>(X, O, gamma, beta, row0, M, D, dp, site, S, mean, rstd, Y);
    AC_LAUNCHED("add_ln_fwd_kernel");
    return AC_OK;
}

constexpr int kLnBwdRows = 32;     // rows per block (8 warps x 4 rows)
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ S,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, int M, int D, Dropout dp, uint32_t site,
                                                         float* __restrict__ dS, float* __restrict__ dO,
                                                         float* __restrict__ partial) {
    __shared__ float s_part[8][2][32 * kLnMaxPerLane];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = D / 32;
    float dg[kLnMaxPerLane], db[kLnMaxPerLane];
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) dg[k] = db[k] = 0.0f;
    for (int i = 0; i < kLnBwdRows / 8; ++i) {
        const int m = blockIdx.x * kLnBwdRows + i * 8 + warp;
        if (m >= M) break;
        const float mu = mean[m], rs = rstd[m];
        float xh[kLnMaxPerLane], gy[kLnMaxPerLane];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < kLnMaxPerLane; ++k) {
            if (k < per) {
                const int d = lane + 32 * k;
                const uint64_t idx = (uint64_t)m * D + d;
                const float dy = dY[idx];
                xh[k] = (S[idx] - mu) * rs;
                gy[k] = dy * gamma[d];
                s1 += gy[k]; s2 += gy[k] * xh[k];
                dg[k] += dy * xh[k]; db[k] += dy;
            }
        }
        s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
#pragma unroll
        for (int k = 0; k < kLnMaxPerLane; ++k) {
            if (k < per) {
                const uint64_t idx = (uint64_t)m * D + lane + 32 * k;
                const float ds = rs * (gy[k] - s1 - xh[k] * s2);
                dS[idx] = ds;
                if (dO != nullptr) dO[idx] = ds * drop_scale(dp.seed, site, idx, dp.p);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kLnMaxPerLane; ++k) { s_part[warp][0][lane + 32 * k] = dg[k]; s_part[warp][1][lane + 32 * k] = db[k]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += 256) {
        const int which = i / D, d = i % D;
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_part[w][which][d];
        partial[(size_t)blockIdx.x * 2 * D + i] = t;
    }
}
size_t ln_bwd_scratch_floats(int M, int D) { return align_up((size_t)cdiv(M, kLnBwdRows) * 2 * D, 32); }
int add_ln_bwd(const float* dY, const float* S, const float* mean, const float* rstd, const float* gamma, int M, int D,
               Dropout dp, uint32_t site, float* dS, float* dO, float* dgamma, float* dbeta, float* scratch, cudaStream_t st) {
    AC_REQUIRE(D % 32 == 0 && D <= 32 * kLnMaxPerLane, "add_ln_bwd: width %d not supported", D);
    if (M <= 0) return AC_OK;
    const int nblk = cdiv(M, kLnBwdRows);
    add_ln_bwd_kernel<<<nblk, 256, 0, st>>>(dY, S, mean, rstd, gamma, M, D, dp, site, dS, dO, scratch);
    AC_LAUNCHED("add_ln_bwd_kernel");
    int rc = AC_OK;
    if (dgamma != nullptr) { rc = colsum(scratch, nblk, D, 2 * D, dgamma, st); if (rc) return rc; }
    if (dbeta != nullptr) { rc = colsum(scratch + D, nblk, D, 2 * D, dbeta, st); if (rc) return rc; }
    return AC_OK;
}

// ------------------------------------------------------------------------------------ elementwise
__global__ void dropout_apply_kernel(float* __restrict__ X, int64_t i0, int64_t n, Dropout dp, uint32_t site) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) X[i0 + i] *= drop_scale(dp.seed, site, (uint64_t)(i0 + i), dp.p);
}
int dropout_apply(float* X, int64_t i0, int64_t n, Dropout dp, uint32_t site, cudaStream_t st) {
    if (n <= 0 || dp.p <= 0.0f) return AC_OK;
    dropout_apply_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(X, i0, n, dp, site);
    AC_LAUNCHED("dropout_apply_kernel");
    return AC_OK;
}
__global__ void relu_drop_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, int64_t n, Dropout dp,
                                     uint32_t site, float* __restrict__ dX) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dX[i] = Y[i] > 0.0f ? dY[i] * drop_scale(dp.seed, site, (uint64_t)i, dp.p) : 0.0f;
}
int relu_drop_bwd(const float* dY, const float* Y, int64_t n, Dropout dp, uint32_t site, float* dX, cudaStream_t st) {
    if (n <= 0) return AC_OK;
    relu_drop_bwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(dY, Y, n, dp, site, dX);
    AC_LAUNCHED("relu_drop_bwd_kernel");
    return AC_OK;
}
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows, int D, float* __restrict__ dst) {
    const int i = blockIdx.x, r = rows[i];
    for (int d = threadIdx.x; d < D; d += blockDim.x) dst[(size_t)i * D + d] = src[(size_t)r * D + d];
}
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows, int D, float* __restrict__ dst) {
    const int i = blockIdx.x, r = rows[i];
    for (int d = threadIdx.x; d < D; d += blockDim.x) dst[(size_t)r * D + d] = src[(size_t)i * D + d];
}
int gather_rows(const float* src, const int* rows, int n, int D, float* dst, cudaStream_t st) {
    if (n <= 0) return AC_OK;
    gather_rows_kernel<<<n, 256, 0, st>>>(src, rows, D, dst);
    AC_LAUNCHED("gather_rows_kernel");
    return AC_OK;
}
int scatter_rows(const float* src, const int* rows, int n, int D, float* dst, cudaStream_t st) {
    if (n <= 0) return AC_OK;
    scatter_rows_kernel<<<n, 256, 0, st>>>(src, rows, D, dst);
    AC_LAUNCHED("scatter_rows_kernel");
    return AC_OK;
}

// deterministic block reductions (256 threads)
__device__ __forceinline__ float block_sum256(float v, float* s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) s[warp] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i];
    __syncthreads();
    return t;
}
__device__ __forceinline__ float block_max256(float v, float* s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) s[warp] = v;
    __syncthreads();
    float t = s[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) t = fmaxf(t, s[i]);
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ X, int ld, int V, int64_t* __restrict__ idx,
                                                          float* __restrict__ logprob) {
    __shared__ float s_f[8];
    __shared__ int s_i[8];
    const float* row = X + (size_t)blockIdx.x * ld;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += 256) {
        const float x = row[v];
        if (x > best) { best = x; bi = v; }          // strided ascending: first maximum per thread
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_f[warp] = best; s_i[warp] = bi; }
    __syncthreads();
    best = s_f[0]; bi = s_i[0];
#pragma unroll
    for (int w = 1; w < 8; ++w)
        if (s_f[w] > best || (s_f[w] == best && s_i[w] < bi)) { best = s_f[w]; bi = s_i[w]; }
    __syncthreads();
    if (logprob != nullptr) {
        float se = 0.0f;
        for (int v = threadIdx.x; v < V; v += 256) se += expf(row[v] - best);
        se = block_sum256(se, s_f);
        if (threadIdx.x == 0) logprob[blockIdx.x] = -logf(se);
    }
    if (threadIdx.x == 0) idx[blockIdx.x] = bi;
}
int argmax_rows(const float* X, int ld, int M, int V, int64_t* idx, float* logprob, cudaStream_t st) {
    if (M <= 0) return AC_OK;
    argmax_rows_kernel<<<M, 256, 0, st>>>(X, ld, V, idx, logprob);
    AC_LAUNCHED("argmax_rows_kernel");
    return AC_OK;
}

// ------------------------------------------------------------------------------------ attention
constexpr int kAttnLd = kAttnHeadDim + 1;   // padded smem rows: conflict-free for both row- and column-wise walks

__device__ __forceinline__ bool attn_masked(const AttnArgs& a, int seq, int i, int j, int kvlen) {
    if (a.causal && j > i) return true;
    if (j >= kvlen) return true;
    return a.key_pad != nullptr && a.key_pad[(size_t)seq * a.Lk + j] != 0;
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnArgs a) {
    extern __shared__ float sm[];
    const int L = a.L, Lk = a.Lk;
    float* sQ = sm; float* sK = sQ + L * kAttnLd; float* sV = sK + Lk * kAttnLd; float* sP = sV + Lk * kAttnLd;   // sP [L][Lk+1]
    const int seq = a.seq0 + blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int kvs = seq % a.n_kv_seq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kvlen = a.kv_len != nullptr ? (int)min((int64_t)Lk, max((int64_t)0, a.kv_len[kvs])) : Lk;
    for (int i = tid; i < L * kAttnHeadDim; i += 128) {
        const int r = i / kAttnHeadDim, d = i % kAttnHeadDim;
        sQ[r * kAttnLd + d] = a.Q[((size_t)seq * L + r) * a.ldq + h * kAttnHeadDim + d];
    }
    for (int i = tid; i < Lk * kAttnHeadDim; i += 128) {
        const int r = i / kAttnHeadDim, d = i % kAttnHeadDim;
        const size_t o = ((size_t)kvs * Lk + r) * a.ldkv + h * kAttnHeadDim + d;
        sK[r * kAttnLd + d] = a.K[o]; sV[r * kAttnLd + d] = a.V[o];
    }
    __syncthreads();
    for (int e = tid; e < L * Lk; e += 128) {
        const int i = e / Lk, j = e % Lk;
        float s = -INFINITY;
        if (!attn_masked(a, seq, i, j, kvlen)) {
            s = 0.0f;
#pragma unroll 16
            for (int d = 0; d < kAttnHeadDim; ++d) s = fmaf(sQ[i * kAttnLd + d], sK[j * kAttnLd + d], s);
            s *= 0.125f;                         // 1 / sqrt(64)
        }
        sP[i * (Lk + 1) + j] = s;
    }
    __syncthreads();
    float* Pg = a.P + ((size_t)seq * a.H + h) * L * Lk;
    for (int i = warp; i < L; i += 4) {
        float* row = sP + i * (Lk + 1);
        float m = -INFINITY;
        for (int j = lane; j < Lk; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float se = 0.0f;
        for (int j = lane; j < Lk; j += 32) { const float e = row[j] == -INFINITY ? 0.0f : expf(row[j] - m); row[j] = e; se += e; }
        se = warp_sum(se);
        const float inv = se > 0.0f ? 1.0f / se : 0.0f;     // a fully masked row (no valid key) attends to nothing
        for (int j = lane; j < Lk; j += 32) {
            const float p = row[j] * inv;
            Pg[(size_t)i * Lk + j] = p;
            row[j] = p * drop_scale(a.dp.seed, a.site, (((uint64_t)seq * a.H + h) * L + i) * Lk + j, a.dp.p);
        }
    }
    __syncthreads();
    for (int e = tid; e < L * kAttnHeadDim; e += 128) {
        const int i = e / kAttnHeadDim, d = e % kAttnHeadDim;
        float o = 0.0f;
        const float* row = sP + i * (Lk + 1);
        for (int j = 0; j < Lk; ++j) o = fmaf(row[j], sV[j * kAttnLd + d], o);
        a.O[((size_t)seq * L + i) * a.ldo + h * kAttnHeadDim + d] = o;
    }
}
static size_t attn_fwd_smem(int L, int Lk) { return ((size_t)(L + 2 * Lk) * kAttnLd + (size_t)L * (Lk + 1)) * sizeof(float); }
int attn_fwd(const AttnArgs& a, cudaStream_t st) {
    AC_REQUIRE(a.L >= 1 && a.L <= kAttnMaxL && a.Lk >= 1 && a.Lk <= kAttnMaxLk, "attn_fwd: L=%d (<= %d) Lk=%d (<= %d)", a.L,
               kAttnMaxL, a.Lk, kAttnMaxLk);
    if (a.n_seq <= 0) return AC_OK;
    static cudaError_t attr_rc = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)attn_fwd_smem(kAttnMaxL, kAttnMaxLk));
    AC_CUDA(attr_rc);
    attn_fwd_kernel<<<a.n_seq * a.H, 128, attn_fwd_smem(a.L, a.Lk), st>>>(a);
    AC_LAUNCHED("attn_fwd_kernel");
    return AC_OK;
}

// One CTA per (kv sequence, head); loops over the query sequences sharing that kv sequence so that dK / dV accumulate
// without atomics (first sequence writes, later ones add: same thread, same address).
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnBwdArgs b) {
    extern __shared__ float sm[];
    const AttnArgs& a = b.f;
    const int L = a.L, Lk = a.Lk, ldp = Lk + 1;
    float* sQ = sm; float* sdO = sQ + L * kAttnLd; float* sK = sdO + L * kAttnLd; float* sV = sK + Lk * kAttnLd;
    float* sP = sV + Lk * kAttnLd; float* sdS = sP + L * ldp; float* sPd = sdS + L * ldp;
    const int kvs = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < Lk * kAttnHeadDim; i += 256) {
        const int r = i / kAttnHeadDim, d = i % kAttnHeadDim;
        const size_t o = ((size_t)kvs * Lk + r) * a.ldkv + h * kAttnHeadDim + d;
        sK[r * kAttnLd + d] = a.K[o]; sV[r * kAttnLd + d] = a.V[o];
    }
    const int reps = a.n_seq / a.n_kv_seq;
    for (int rep = 0; rep < reps; ++rep) {
        const int seq = rep * a.n_kv_seq + kvs;
        __syncthreads();
        for (int i = tid; i < L * kAttnHeadDim; i += 256) {
            const int r = i / kAttnHeadDim, d = i % kAttnHeadDim;
            sQ[r * kAttnLd + d] = a.Q[((size_t)seq * L + r) * a.ldq + h * kAttnHeadDim + d];
            sdO[r * kAttnLd + d] = b.dO[((size_t)seq * L + r) * b.lddo + h * kAttnHeadDim + d];
        }
        const float* Pg = a.P + ((size_t)seq * a.H + h) * L * Lk;
        for (int e = tid; e < L * Lk; e += 256) sP[(e / Lk) * ldp + e % Lk] = Pg[e];
        __syncthreads();
        // dPd = dO V^T;  dP = dPd * mask;  Pd = P * mask
        for (int e = tid; e < L * Lk; e += 256) {
            const int i = e / Lk, j = e % Lk;
            const float p = sP[i * ldp + j];
            float g = 0.0f;
            if (p != 0.0f) {
#pragma unroll 16
                for (int d = 0; d < kAttnHeadDim; ++d) g = fmaf(sdO[i * kAttnLd + d], sV[j * kAttnLd + d], g);
            }
            const float ds = drop_scale(a.dp.seed, a.site, (((uint64_t)seq * a.H + h) * L + i) * Lk + j, a.dp.p);
            sdS[i * ldp + j] = g * ds;        // dP for now
            sPd[i * ldp + j] = p * ds;
        }
        __syncthreads();
        for (int i = warp; i < L; i += 8) {
            float dot = 0.0f;
            for (int j = lane; j < Lk; j += 32) dot += sdS[i * ldp + j] * sP[i * ldp + j];
            dot = warp_sum(dot);
            for (int j = lane; j < Lk; j += 32) sdS[i * ldp + j] = sP[i * ldp + j] * (sdS[i * ldp + j] - dot);
        }
        __syncthreads();
        for (int e = tid; e < L * kAttnHeadDim; e += 256) {          // dQ = dS K / 8
            const int i = e / kAttnHeadDim, d = e % kAttnHeadDim;
            float g = 0.0f;
            for (int j = 0; j < Lk; ++j) g = fmaf(sdS[i * ldp + j], sK[j * kAttnLd + d], g);
            b.dQ[((size_t)seq * L + i) * b.lddq + h * kAttnHeadDim + d] = g * 0.125f;
        }
        for (int e = tid; e < Lk * kAttnHeadDim; e += 256) {         // dK = dS^T Q / 8, dV = Pd^T dO
            const int j = e / kAttnHeadDim, d = e % kAttnHeadDim;
            float gk = 0.0f, gv = 0.0f;
            for (int i = 0; i < L; ++i) {
                gk = fmaf(sdS[i * ldp + j], sQ[i * kAttnLd + d], gk);
                gv = fmaf(sPd[i * ldp + j], sdO[i * kAttnLd + d], gv);
            }
            const size_t o = ((size_t)kvs * Lk + j) * b.lddkv + h * kAttnHeadDim + d;
            if (rep == 0) { b.dK[o] = gk * 0.125f; b.dV[o] = gv; }
            else { b.dK[o] += gk * 0.125f; b.dV[o] += gv; }
        }
    }
}
static size_t attn_bwd_smem(int L, int Lk) { return ((size_t)(2 * L + 2 * Lk) * kAttnLd + (size_t)3 * L * (Lk + 1)) * sizeof(float); }
int attn_bwd(const AttnBwdArgs& b, cudaStream_t st) {
    const AttnArgs& a = b.f;
    AC_REQUIRE(a.L >= 1 && a.L <= kAttnMaxLBwd && a.Lk >= 1 && a.Lk <= kAttnMaxLk, "attn_bwd: L=%d (<= %d) Lk=%d (<= %d)", a.L,
               kAttnMaxLBwd, a.Lk, kAttnMaxLk);
    AC_REQUIRE(a.seq0 == 0 && a.n_kv_seq > 0 && a.n_seq % a.n_kv_seq == 0, "attn_bwd: sequences must be a multiple of kv sequences");
    if (a.n_seq <= 0) return AC_OK;
    static cudaError_t attr_rc = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)attn_bwd_smem(kAttnMaxLBwd, kAttnMaxLk));
    AC_CUDA(attr_rc);
    attn_bwd_kernel<<<a.n_kv_seq * a.H, 256, attn_bwd_smem(a.L, a.Lk), st>>>(b);
    AC_LAUNCHED("attn_bwd_kernel");
    return AC_OK;
}

// ------------------------------------------------------------------------------------ label-smoothing cross entropy
// loss.py:51-74: preds = log_softmax(logit); true_dist = s / (V - 1) off target, 1 - s on target;
// loss = sum_m mask_m * (-sum_v true_dist * preds) / sum_m mask_m,  mask[b, t] = t < tgt_len[b].
__global__ void __launch_bounds__(256) ls_ce_kernel(const float* __restrict__ logit, int ldl, const int64_t* __restrict__ tgt, int ld_tgt,
                                                    const int64_t* __restrict__ tgt_len, int B, int L, int V, float smoothing,
                                                    float grad_scale, float* __restrict__ row_loss, float* __restrict__ dlogit) {
    __shared__ float s_f[8];
    __shared__ float s_ntok;
    const int m = blockIdx.x, b = m / L, t = m % L;
    const float* row = logit + (size_t)m * ldl;
    if (threadIdx.x == 0) {
        int64_t n = 0;
        for (int i = 0; i < B; ++i) n += min((int64_t)L, max((int64_t)0, tgt_len[i]));
        s_ntok = (float)n;
    }
    const bool valid = t < tgt_len[b];
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += 256) mx = fmaxf(mx, row[v]);
    mx = block_max256(mx, s_f);
    float se = 0.0f, sl = 0.0f;
    for (int v = threadIdx.x; v < V; v += 256) { const float x = row[v]; se += expf(x - mx); sl += x; }
    se = block_sum256(se, s_f);
    sl = block_sum256(sl, s_f);
    const float lse = mx + logf(se);
    const int64_t tg = min((int64_t)V - 1, max((int64_t)0, tgt[(size_t)b * ld_tgt + t]));
    const float eps = smoothing / (float)(V - 1), conf = 1.0f - smoothing;
    if (threadIdx.x == 0) {
        const float lp_t = row[tg] - lse;
        const float sum_lp = sl - (float)V * lse;
        row_loss[m] = valid ? -(conf * lp_t + eps * (sum_lp - lp_t)) : 0.0f;
    }
    if (dlogit != nullptr) {
        const float w = valid ? grad_scale / s_ntok : 0.0f;
        float* drow = dlogit + (size_t)m * ldl;
        for (int v = threadIdx.x; v < ldl; v += 256)
            drow[v] = v < V ? w * (expf(row[v] - lse) - (v == tg ? conf : eps)) : 0.0f;
    }
}
__global__ void __launch_bounds__(256) ls_ce_reduce_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ tgt_len,
                                                           int B, int L, float* __restrict__ loss) {
    __shared__ float s_f[8];
    float acc = 0.0f;
    for (int m = threadIdx.x; m < B * L; m += 256) acc += row_loss[m];
    acc = block_sum256(acc, s_f);
    if (threadIdx.x == 0) {
        int64_t n = 0;
        for (int i = 0; i < B; ++i) n += min((int64_t)L, max((int64_t)0, tgt_len[i]));
        loss[0] = acc / (float)n;
    }
}

// ------------------------------------------------------------------------------------ clip_grad_norm_ + Adam
constexpr int kNormBlocks = 592;     // 4 per SM
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float scale, float* __restrict__ partial) {
    __shared__ float s_f[8];
    float acc = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float x = g[i] * scale;
        acc = fmaf(x, x, acc);
    }
    acc = block_sum256(acc, s_f);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
struct AdamArgs {
    float* p; const float* g; float* m; float* v; int64_t n;
    float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
    const float* partial; int n_partial;
    const float* loss;       // nullable: a NaN loss skips the update (run.py:123)
    int* step;               // device step counter (incremented by the kernel when the update is applied)
    float* norm_out;         // nullable: total gradient norm before clipping
};
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    __shared__ float s_clip;
    __shared__ double s_bc1, s_bc2;
    __shared__ int s_skip;
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int i = 0; i < a.n_partial; ++i) tot += a.partial[i];      // fixed order: every block computes the same value
        const float norm = sqrtf(tot);
        s_clip = a.max_norm > 0.0f ? fminf(1.0f, a.max_norm / (norm + 1e-6f)) : 1.0f;
        const int step = *a.step + 1;
        s_bc1 = 1.0 - pow((double)a.beta1, (double)step);
        s_bc2 = 1.0 - pow((double)a.beta2, (double)step);
        s_skip = (a.loss != nullptr && isnan(*a.loss)) || isnan(norm) ? 1 : 0;
        if (blockIdx.x == 0 && a.norm_out != nullptr) *a.norm_out = norm;
    }
    __syncthreads();
    if (s_skip) return;
    const float gs = a.grad_scale * s_clip;
    const float step_size = (float)((double)a.lr / s_bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(s_bc2));
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * 256) {
        const float p = a.p[i];
        const float g = fmaf(a.weight_decay, p, a.g[i] * gs);
        const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
        a.m[i] = m; a.v[i] = v;
        a.p[i] = p - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    }
}
__global__ void adam_step_kernel(int* step, const float* loss, const float* partial, int n_partial) {
    float tot = 0.0f;
    for (int i = 0; i < n_partial; ++i) tot += partial[i];
    if (!((loss != nullptr && isnan(*loss)) || isnan(tot))) *step += 1;
}

__global__ void specaug_kernel(float* __restrict__ x, int F, int T, const int* __restrict__ stripes, int ns, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T), f = (int)((i / T) % F), b = (int)(i / ((int64_t)T * F));
    const int* s = stripes + (size_t)b * 4 * ns;
    bool drop = false;
    for (int k = 0; k < ns; ++k) {
        drop = drop || (t >= s[2 * k] && t < s[2 * k] + s[2 * k + 1]);
        drop = drop || (f >= s[2 * (ns + k)] && f < s[2 * (ns + k)] + s[2 * (ns + k) + 1]);
    }
    if (drop) x[i] = 0.0f;
}

}  // namespace ac

extern "C" {

int ac_ls_ce_fwd_bwd(const float* logit_dev, int ld_logit, const int64_t* tgt_dev, int ld_tgt, const int64_t* tgt_len_dev, int B, int L, int V,
                     float smoothing, float grad_scale, float* loss_dev, float* dlogit_dev, void* workspace_dev,
                     size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 1 && L >= 1 && V >= 2 && ld_logit >= V && logit_dev && tgt_dev && tgt_len_dev && loss_dev, "ac_ls_ce_fwd_bwd: bad argument");
    AC_REQUIRE(workspace_dev && workspace_bytes >= (size_t)B * L * sizeof(float), "ac_ls_ce_fwd_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* row_loss = (float*)workspace_dev;
    AC_TIMED("ls_ce", st);
    ls_ce_kernel<<<B * L, 256, 0, st>>>(logit_dev, ld_logit, tgt_dev, ld_tgt, tgt_len_dev, B, L, V, smoothing, grad_scale, row_loss, dlogit_dev);
    AC_LAUNCHED("ls_ce_kernel");
    ls_ce_reduce_kernel<<<1, 256, 0, st>>>(row_loss, tgt_len_dev, B, L, loss_dev);
    AC_LAUNCHED("ls_ce_reduce_kernel");
    return AC_OK;
}

// SpecAugment (cnn_encoder.py:352-353,424-425: torchlibrosa SpecAugmentation, training only): zero `n_stripes` time stripes
// and `n_stripes` mel stripes per clip of the dB log-mel.  stripes_dev [batch][2 * n_stripes][2] int32 = (begin, width):
// the first n_stripes entries are frame ranges, the rest mel ranges (drawn on the host by the module, in the library's
// draw order).  lms_dev [batch, n_mels, n_frames], in place.
int ac_specaug_apply(float* lms_dev, int batch, int n_mels, int n_frames, const int* stripes_dev, int n_stripes, void* stream) {
    using namespace ac;
    AC_REQUIRE(batch >= 0 && n_mels >= 1 && n_frames >= 1 && n_stripes >= 0 && n_stripes <= 8, "ac_specaug_apply: bad argument");
    if (batch == 0 || n_stripes == 0) return AC_OK;
    AC_REQUIRE(lms_dev && stripes_dev, "ac_specaug_apply: null argument");
    const int64_t total = (int64_t)batch * n_mels * n_frames;
    specaug_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(lms_dev, n_mels, n_frames, stripes_dev, n_stripes, total);
    AC_LAUNCHED("specaug_kernel");
    return AC_OK;
}

// sample_next_word(method="greedy") of captioning/models/base.py:214-218 over a batch of logit rows:
// idx_dev[m] = first arg-max of logit_dev[m, :V] (row stride ld), logprob_dev[m] (nullable) = its log-softmax value.
int ac_argmax_rows(const float* logit_dev, int ld, int M, int V, int64_t* idx_dev, float* logprob_dev, void* stream) {
    using namespace ac;
    AC_REQUIRE(logit_dev && idx_dev && ld >= V && V >= 1 && M >= 0, "ac_argmax_rows: bad argument");
    return argmax_rows(logit_dev, ld, M, V, idx_dev, logprob_dev, (cudaStream_t)stream);
}

size_t ac_clip_adam_workspace_bytes(void) { return ac::kNormBlocks * sizeof(float); }

int ac_clip_adam(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                 const float* loss_dev, int* step_dev, float* norm_out_dev, void* workspace_dev, size_t workspace_bytes,
                 void* stream) {
    using namespace ac;
    AC_REQUIRE(param_dev && grad_dev && exp_avg_dev && exp_avg_sq_dev && step_dev && n > 0, "ac_clip_adam: bad argument");
    AC_REQUIRE(workspace_dev && workspace_bytes >= ac_clip_adam_workspace_bytes(), "ac_clip_adam: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)workspace_dev;
    AC_TIMED("clip_adam", st);
    sumsq_kernel<<<kNormBlocks, 256, 0, st>>>(grad_dev, n, grad_scale, partial);
    AC_LAUNCHED("sumsq_kernel");
    AdamArgs a;
    a.p = param_dev; a.g = grad_dev; a.m = exp_avg_dev; a.v = exp_avg_sq_dev; a.n = n; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2;
    a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm; a.grad_scale = grad_scale; a.partial = partial;
    a.n_partial = kNormBlocks; a.loss = loss_dev; a.step = step_dev; a.norm_out = norm_out_dev;
    adam_kernel<<<kNormBlocks, 256, 0, st>>>(a);
    AC_LAUNCHED("adam_kernel");
    adam_step_kernel<<<1, 1, 0, st>>>(step_dev, loss_dev, partial, kNormBlocks);
    AC_LAUNCHED("adam_step_kernel");
    return AC_OK;
}

}  // extern "C"
