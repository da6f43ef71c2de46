// EfficientNet-B2 feature extractor (eval mode), fp32, NHWC.
//
// Replaces `_EffiNet.forward` of the reference (captioning/models/hf_wrapper.py:218-241:
// efficientnet_pytorch 0.7.1 `extract_features` on a [B,1,F,T] log-mel + mean over F).
// The block plan (23 MBConv blocks, "static same" padding computed for a 260x260 image and
// applied to the 64 x T spectrogram) is rebuilt here from the B2 coefficients exactly as the
// package does; tests compare it with the oracle's plan through ac_effb2_block_info.
//
// Data layout in HBM: activations [B, H(freq), W(time), C] fp32 (channels innermost), so a
// 1x1 convolution is the TN GEMM of gemm.cu and the depthwise convolution vectorises over
// channels.  BatchNorm is folded to a per-channel scale/bias at pack time and applied in the
// producing kernel's epilogue, together with swish, the SE gate (applied while the projection
// GEMM loads its A operand) and the residual add.  The top_db clamp of AmplitudeToDB is
// applied while the stem convolution loads the spectrogram.
#include <math.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"
#include "gemm.cuh"

namespace ac {

struct BlockPlan {
    int cin, cout, expand, k, s, pad_lo, pad_hi, nsq, skip;
    int cexp() const { return cin * expand; }
};

static int round_filters(int f) {
    double x = f * 1.1;
    int nf = std::max(8, (int)(x + 4.0) / 8 * 8);
    if (nf < 0.9 * x) nf += 8;
    return nf;
}
static int round_repeats(int r) { return (int)ceil(1.2 * r); }

struct Plan {
    int stem_out, stem_pad_lo, stem_pad_hi, head_in, head_out;
    std::vector<BlockPlan> blocks;
    Plan() {
        static const int base[7][6] = {  // repeats, k, s, expand, in, out   (efficientnet-b0 stages)
            {1, 3, 1, 1, 32, 16},  {2, 3, 2, 6, 16, 24},  {2, 5, 2, 6, 24, 40}, {3, 3, 2, 6, 40, 80},
            {3, 5, 1, 6, 80, 112}, {4, 5, 2, 6, 112, 192}, {1, 3, 1, 6, 192, 320}};
        auto same_pad = [](int img, int k, int s, int& lo, int& hi) {
            int o = (img + s - 1) / s;
            int p = std::max((o - 1) * s + (k - 1) + 1 - img, 0);
            lo = p / 2; hi = p - p / 2;
        };
        int img = 260;
        stem_out = round_filters(32);
        same_pad(img, 3, 2, stem_pad_lo, stem_pad_hi);
        img = (img + 1) / 2;
        for (auto& st : base) {
            int cin = round_filters(st[4]), cout = round_filters(st[5]), rep = round_repeats(st[0]);
            for (int r = 0; r < rep; ++r) {
                BlockPlan b;
                b.cin = r == 0 ? cin : cout; b.cout = cout; b.expand = st[3]; b.k = st[1];
                b.s = r == 0 ? st[2] : 1;
                same_pad(img, b.k, b.s, b.pad_lo, b.pad_hi);
                b.nsq = std::max(1, (int)(b.cin * 0.25));
                b.skip = (r > 0 && b.cin == b.cout) ? 1 : 0;   // first block of a stage never skips
                blocks.push_back(b);
                img = (img + b.s - 1) / b.s;
            }
            head_in = cout;
        }
        head_out = round_filters(1280);
    }
};

static const Plan& plan() { static Plan p; return p; }

// ----------------------------------------------------------------------------- pack kernels
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ bias, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float s = gamma[i] / sqrtf(var[i] + eps);
        scale[i] = s;
        bias[i] = beta[i] - mean[i] * s;
    }
}
// [C][KK] -> [KK][C]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) {
        int r = i / cols, c = i % cols;
        out[(size_t)c * rows + r] = in[i];
    }
}

// ----------------------------------------------------------------------------- stem
// lms [B, H, W] -> out [B, Ho, Wo, C = 32]; 3x3 stride 2, BN + swish.  One CTA = kStemCols consecutive output columns
// of one output row: the 3 x (2 kStemCols + 1) input patch is staged in shared memory with coalesced loads (dB floor
// applied on the way in, zero padding left at zero), then thread = 4 channels x kStemPx columns with the 9 weight quads
// in registers.  (The one-pixel-per-thread version issued 18 global loads per 128-bit store and sat at 22 % occupancy
// waiting for them: 107 us for the 131 MB it writes.)
constexpr int kStemPx = 4;
constexpr int kStemCols = 128;                 // (256 threads / 8 channel quads) x kStemPx
__global__ void __launch_bounds__(256, 3)
stem_kernel(const float* __restrict__ lms, const float* __restrict__ gmax, float top_db,
            const float* __restrict__ w /*[9][C]*/, const float* __restrict__ scale,
            const float* __restrict__ bias, float* __restrict__ out, int H, int W, int Ho, int Wo,
            int pad_lo, int tiles_w) {
    constexpr int NP = 2 * kStemCols + 1;
    __shared__ float s_in[3][NP + 3];
    const int tile = blockIdx.x % tiles_w;
    const int ho = (blockIdx.x / tiles_w) % Ho;
    const int b = blockIdx.x / (tiles_w * Ho);
    const int tid = threadIdx.x;
    const float floor_v = gmax ? (*gmax - top_db) : -INFINITY;
    const float* x = lms + (size_t)b * H * W;
    const int wo_base = tile * kStemCols, iw_base = wo_base * 2 - pad_lo;
    for (int i = tid; i < 3 * NP; i += 256) {
        const int kh = i / NP, j = i - kh * NP;
        const int ih = ho * 2 + kh - pad_lo, iw = iw_base + j;
        s_in[kh][j] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? fmaxf(__ldg(x + (size_t)ih * W + iw), floor_v) : 0.0f;
    }
    const int c4 = tid & 7, g = tid >> 3;
    float4 wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg(reinterpret_cast<const float4*>(w + t * 32) + c4);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    __syncthreads();
    const int wo0 = wo_base + g * kStemPx;
    if (wo0 >= Wo) return;
    float in[3][2 * kStemPx + 1];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < 2 * kStemPx + 1; ++j) in[kh][j] = s_in[kh][g * 2 * kStemPx + j];
    float4* o4 = reinterpret_cast<float4*>(out) + (((size_t)b * Ho + ho) * Wo + wo0) * 8 + c4;
#pragma unroll
    for (int px = 0; px < kStemPx; ++px) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float v = in[kh][2 * px + kw];
                const float4 ww = wt[kh * 3 + kw];
                acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y);
                acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
            }
        if (wo0 + px < Wo)
            o4[(size_t)px * 8] = make_float4(fast_swish(fmaf(acc.x, sc.x, bb.x)), fast_swish(fmaf(acc.y, sc.y, bb.y)),
                                             fast_swish(fmaf(acc.z, sc.z, bb.z)), fast_swish(fmaf(acc.w, sc.w, bb.w)));
    }
}

// ----------------------------------------------------------------------------- squeeze-and-excitation
// (the depthwise convolution itself lives in dwconv_tma.cu)
// One thread-block CLUSTER per clip (P = 1..8 CTAs): every CTA owns 1/P of the channels (mean and gate) and
// 1/P of the squeezed units, so each CTA streams only 1/P of the two FC weight matrices; the mean vector and
// the squeezed vector are exchanged through distributed shared memory between the phases.
// wr [nsq][C], we_t [nsq][C] (transposed at pack time so that both FC layers read coalesced rows).
constexpr int kSeThreads = 256;
__global__ void __launch_bounds__(kSeThreads)
se_kernel(const float* __restrict__ partial, int strips, float inv_hw, const float* __restrict__ wr,
          const float* __restrict__ br, const float* __restrict__ we_t, const float* __restrict__ be,
          float* __restrict__ gate, int C, int nsq) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int P = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) float s_se[];   // mean[C] | r[nsq] | scratch[kSeThreads * 4]
    float* s_mean = s_se;
    float* s_r = s_se + C;
    float4* s_scr = reinterpret_cast<float4*>(s_se + ((C + nsq + 3) / 4) * 4);
    const int b = blockIdx.x / P, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c4n = C / 4;
    const int q0 = (int)((int64_t)c4n * rank / P), q1 = (int)((int64_t)c4n * (rank + 1) / P);   // my channel quads
    const int j0 = (int)((int64_t)nsq * rank / P), j1 = (int)((int64_t)nsq * (rank + 1) / P);   // my squeezed units
    const float4* p4 = reinterpret_cast<const float4*>(partial + (size_t)b * strips * C);
    pdl_trigger();
    pdl_wait();        // `partial` is the depthwise kernel's output
    // 1. channel means of my quads: thread = (strip group g, quad), fixed-order two-level sum (deterministic)
    for (int cbase = q0; cbase < q1; cbase += kSeThreads) {
        const int width = min(q1 - cbase, kSeThreads);
        const int G = max(1, min(kSeThreads / width, strips));
        const int g = tid / width, c4 = cbase + tid % width;
        if (g < G) {
            // four loads in flight per thread (the early blocks have 512 slots per clip: eight dependent L2 round trips
            // otherwise); the partial sums are combined in a fixed order
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
            int q = g;
            for (; q + 3 * G < strips; q += 4 * G) {
                const float4 v0 = __ldg(p4 + (size_t)q * c4n + c4), v1 = __ldg(p4 + (size_t)(q + G) * c4n + c4);
                const float4 v2 = __ldg(p4 + (size_t)(q + 2 * G) * c4n + c4), v3 = __ldg(p4 + (size_t)(q + 3 * G) * c4n + c4);
                s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
                s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
                s2.x += v2.x; s2.y += v2.y; s2.z += v2.z; s2.w += v2.w;
                s3.x += v3.x; s3.y += v3.y; s3.z += v3.z; s3.w += v3.w;
            }
            for (; q < strips; q += G) {
                const float4 v = __ldg(p4 + (size_t)q * c4n + c4);
                s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
            }
            s_scr[tid] = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y),
                                     (s0.z + s1.z) + (s2.z + s3.z), (s0.w + s1.w) + (s2.w + s3.w));
        }
        __syncthreads();
        if (tid < width) {
            float4 t = s_scr[tid];
            for (int q = 1; q < G; ++q) {
                const float4 u = s_scr[q * width + tid];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            reinterpret_cast<float4*>(s_mean)[c4] = make_float4(t.x * inv_hw, t.y * inv_hw, t.z * inv_hw, t.w * inv_hw);
        }
        __syncthreads();
    }
    if (P > 1) {
        cluster.sync();
        for (int c4 = tid; c4 < c4n; c4 += kSeThreads) {        // gather the other CTAs' quads
            if (c4 >= q0 && c4 < q1) continue;
            int owner = (int)(((int64_t)(c4 + 1) * P - 1) / c4n);
            while ((int)((int64_t)c4n * owner / P) > c4) --owner;
            while ((int)((int64_t)c4n * (owner + 1) / P) <= c4) ++owner;
            const float4* remote = reinterpret_cast<const float4*>(cluster.map_shared_rank(s_mean, owner));
            reinterpret_cast<float4*>(s_mean)[c4] = remote[c4];
        }
        __syncthreads();
    }
    // 2. squeeze: r[j] = swish(wr[j] . mean + br[j]) for my units, one warp per unit; the row is read as float4 with
    //    four independent loads in flight per lane (the phase is a chain of L2 round trips, not bandwidth)
    for (int j = j0 + warp; j < j1; j += kSeThreads / 32) {
        const float4* w4 = reinterpret_cast<const float4*>(wr + (size_t)j * C);
        const float4* m4 = reinterpret_cast<const float4*>(s_mean);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int c = lane;
        for (; c + 96 < c4n; c += 128) {
            const float4 a0 = __ldg(w4 + c), a1 = __ldg(w4 + c + 32), a2 = __ldg(w4 + c + 64), a3 = __ldg(w4 + c + 96);
            const float4 b0 = m4[c], b1 = m4[c + 32], b2 = m4[c + 64], b3 = m4[c + 96];
            s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
            s1 = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, s1))));
            s2 = fmaf(a2.x, b2.x, fmaf(a2.y, b2.y, fmaf(a2.z, b2.z, fmaf(a2.w, b2.w, s2))));
            s3 = fmaf(a3.x, b3.x, fmaf(a3.y, b3.y, fmaf(a3.z, b3.z, fmaf(a3.w, b3.w, s3))));
        }
        for (; c < c4n; c += 32) {
            const float4 a0 = __ldg(w4 + c), b0 = m4[c];
            s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
        }
        const float s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) s_r[j] = swishf(s + br[j]);
    }
    if (P > 1) {
        cluster.sync();
        for (int j = tid; j < nsq; j += kSeThreads) {
            if (j >= j0 && j < j1) continue;
            int owner = (int)(((int64_t)(j + 1) * P - 1) / nsq);
            while ((int)((int64_t)nsq * owner / P) > j) --owner;
            while ((int)((int64_t)nsq * (owner + 1) / P) <= j) ++owner;
            s_r[j] = cluster.map_shared_rank(s_r, owner)[j];
        }
    }
    __syncthreads();
    // 3. excite: gate[c] = sigmoid(sum_j we_t[j][c] * r[j] + be[c]) for my channel quads.  thread = (unit group g, quad):
    //    the units are dealt round-robin to G groups so that every thread has only nsq / G independent loads to wait for;
    //    the groups' partial sums meet in shared memory (fixed order: deterministic)
    for (int cbase = q0; cbase < q1; cbase += kSeThreads) {
        const int width = min(q1 - cbase, kSeThreads);
        const int G = max(1, min(kSeThreads / width, nsq));
        const int g = tid / width, c4 = cbase + tid % width;
        if (g < G) {
            const float4* w4 = reinterpret_cast<const float4*>(we_t) + c4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            int j = g;
            for (; j + 3 * G < nsq; j += 4 * G) {
                const float4 a0 = __ldg(w4 + (size_t)j * c4n), a1 = __ldg(w4 + (size_t)(j + G) * c4n);
                const float4 a2 = __ldg(w4 + (size_t)(j + 2 * G) * c4n), a3 = __ldg(w4 + (size_t)(j + 3 * G) * c4n);
                const float r0 = s_r[j], r1 = s_r[j + G], r2 = s_r[j + 2 * G], r3 = s_r[j + 3 * G];
                s.x = fmaf(a3.x, r3, fmaf(a2.x, r2, fmaf(a1.x, r1, fmaf(a0.x, r0, s.x))));
                s.y = fmaf(a3.y, r3, fmaf(a2.y, r2, fmaf(a1.y, r1, fmaf(a0.y, r0, s.y))));
                s.z = fmaf(a3.z, r3, fmaf(a2.z, r2, fmaf(a1.z, r1, fmaf(a0.z, r0, s.z))));
                s.w = fmaf(a3.w, r3, fmaf(a2.w, r2, fmaf(a1.w, r1, fmaf(a0.w, r0, s.w))));
            }
            for (; j < nsq; j += G) {
                const float4 a0 = __ldg(w4 + (size_t)j * c4n);
                const float r0 = s_r[j];
                s.x = fmaf(a0.x, r0, s.x); s.y = fmaf(a0.y, r0, s.y); s.z = fmaf(a0.z, r0, s.z); s.w = fmaf(a0.w, r0, s.w);
            }
            s_scr[tid] = s;
        }
        __syncthreads();
        if (tid < width) {
            float4 t = s_scr[tid];
            for (int q = 1; q < G; ++q) {
                const float4 u = s_scr[q * width + tid];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            const float4 bq = __ldg(reinterpret_cast<const float4*>(be) + c4);
            reinterpret_cast<float4*>(gate + (size_t)b * C)[c4] =
                make_float4(sigmoidf_(t.x + bq.x), sigmoidf_(t.y + bq.y), sigmoidf_(t.z + bq.z), sigmoidf_(t.w + bq.w));
        }
        __syncthreads();
    }
    if (P > 1) cluster.sync();   // nobody exits while a peer may still read its shared memory
}

// The same for the layers whose FC weights are small (blocks 0-16: <= 86 KB per matrix): P plain CTAs per clip, no cluster.  Every
// CTA computes ALL channel means and ALL squeezed units itself (redundant L2 reads of a few KB) and only its own slice of
// the gates, so nothing is exchanged: no cluster barriers (each a MEMBAR.ALL.GPU + CCTL.IVALL), no DSMEM gathers.
__global__ void __launch_bounds__(kSeThreads)
se_solo_kernel(const float* __restrict__ partial, int strips, float inv_hw, const float* __restrict__ wr,
               const float* __restrict__ br, const float* __restrict__ we_t, const float* __restrict__ be,
               float* __restrict__ gate, int C, int nsq, int P) {
    extern __shared__ __align__(16) float s_se[];   // mean[C] | r[nsq] | scratch[kSeThreads * 4]
    float* s_mean = s_se;
    float* s_r = s_se + C;
    float4* s_scr = reinterpret_cast<float4*>(s_se + ((C + nsq + 3) / 4) * 4);
    const int b = blockIdx.x / P, rank = blockIdx.x - b * P, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c4n = C / 4;
    const int q0 = (int)((int64_t)c4n * rank / P), q1 = (int)((int64_t)c4n * (rank + 1) / P);   // my gate quads
    const float4* p4 = reinterpret_cast<const float4*>(partial + (size_t)b * strips * C);
    pdl_trigger();
    pdl_wait();        // `partial` is the depthwise kernel's output
    for (int cbase = 0; cbase < c4n; cbase += kSeThreads) {
        const int width = min(c4n - cbase, kSeThreads);
        const int G = max(1, min(kSeThreads / width, strips));
        const int g = tid / width, c4 = cbase + tid % width;
        if (g < G) {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
            int q = g;
            for (; q + 3 * G < strips; q += 4 * G) {
                const float4 v0 = __ldg(p4 + (size_t)q * c4n + c4), v1 = __ldg(p4 + (size_t)(q + G) * c4n + c4);
                const float4 v2 = __ldg(p4 + (size_t)(q + 2 * G) * c4n + c4), v3 = __ldg(p4 + (size_t)(q + 3 * G) * c4n + c4);
                s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
                s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
                s2.x += v2.x; s2.y += v2.y; s2.z += v2.z; s2.w += v2.w;
                s3.x += v3.x; s3.y += v3.y; s3.z += v3.z; s3.w += v3.w;
            }
            for (; q < strips; q += G) {
                const float4 v = __ldg(p4 + (size_t)q * c4n + c4);
                s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
            }
            s_scr[tid] = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y),
                                     (s0.z + s1.z) + (s2.z + s3.z), (s0.w + s1.w) + (s2.w + s3.w));
        }
        __syncthreads();
        if (tid < width) {
            float4 t = s_scr[tid];
            for (int q = 1; q < G; ++q) {
                const float4 u = s_scr[q * width + tid];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            reinterpret_cast<float4*>(s_mean)[c4] = make_float4(t.x * inv_hw, t.y * inv_hw, t.z * inv_hw, t.w * inv_hw);
        }
        __syncthreads();
    }
    for (int j = warp; j < nsq; j += kSeThreads / 32) {
        const float4* w4 = reinterpret_cast<const float4*>(wr + (size_t)j * C);
        const float4* m4 = reinterpret_cast<const float4*>(s_mean);
        float s0 = 0.f, s1 = 0.f;
        int c = lane;
        for (; c + 32 < c4n; c += 64) {
            const float4 a0 = __ldg(w4 + c), a1 = __ldg(w4 + c + 32);
            const float4 b0 = m4[c], b1 = m4[c + 32];
            s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
            s1 = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, s1))));
        }
        for (; c < c4n; c += 32) {
            const float4 a0 = __ldg(w4 + c), b0 = m4[c];
            s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
        }
        const float s = warp_sum(s0 + s1);
        if (lane == 0) s_r[j] = swishf(s + br[j]);
    }
    __syncthreads();
    for (int cbase = q0; cbase < q1; cbase += kSeThreads) {
        const int width = min(q1 - cbase, kSeThreads);
        const int G = max(1, min(kSeThreads / width, nsq));
        const int g = tid / width, c4 = cbase + tid % width;
        if (g < G) {
            const float4* w4 = reinterpret_cast<const float4*>(we_t) + c4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = g; j < nsq; j += G) {
                const float4 a0 = __ldg(w4 + (size_t)j * c4n);
                const float r0 = s_r[j];
                s.x = fmaf(a0.x, r0, s.x); s.y = fmaf(a0.y, r0, s.y); s.z = fmaf(a0.z, r0, s.z); s.w = fmaf(a0.w, r0, s.w);
            }
            s_scr[tid] = s;
        }
        __syncthreads();
        if (tid < width) {
            float4 t = s_scr[tid];
            for (int q = 1; q < G; ++q) {
                const float4 u = s_scr[q * width + tid];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            const float4 bq = __ldg(reinterpret_cast<const float4*>(be) + c4);
            reinterpret_cast<float4*>(gate + (size_t)b * C)[c4] =
                make_float4(sigmoidf_(t.x + bq.x), sigmoidf_(t.y + bq.y), sigmoidf_(t.z + bq.z), sigmoidf_(t.w + bq.w));
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------- head tail
// y [B, H, W, C] -> out [B, W, C] = mean over H  ('b c f t -> b t c', 'mean')
__global__ void freq_mean_kernel(const float* __restrict__ y, float* __restrict__ out, int H, int W, int C,
                                 int64_t total /* B*W*C/4 */) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4n = C / 4;
    int c4 = (int)(idx % c4n);
    int64_t p = idx / c4n;
    int w = (int)(p % W);
    int b = (int)(p / W);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = 0; h < H; ++h) {
        float4 v = __ldg(y4 + (((size_t)b * H + h) * W + w) * c4n + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.0f / (float)H;
    reinterpret_cast<float4*>(out)[idx] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

// x [B,T,D], lens [B] -> out [B,D] = sum_{t<len} x / len
__global__ void masked_mean_kernel(const float* __restrict__ x, const int64_t* __restrict__ lens, int T, int D,
                                   float* __restrict__ out) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int64_t len = lens[b];
    const int n = (int)min((int64_t)T, max((int64_t)0, len));
    float s = 0.f;
    for (int t = 0; t < n; ++t) s += x[((size_t)b * T + t) * D + d];
    out[(size_t)b * D + d] = s / (float)len;
}

constexpr int kNarrowLevels = 2;                       // extra tensor-core images per 1x1 convolution: n-tiles <= 64, <= 32 columns
constexpr int kNarrowCap[kNarrowLevels] = {64, 32};
struct ConvBN {
    float* w; float* scale; float* bias; TcWeight tw;
    TcWeight narrow[kNarrowLevels];
    // A launch with fewer than 64 tiles (small batches; at 64 clips no layer) takes the next narrower image: B = 32 encoder
    // GEMM time 1.19 -> 1.13 ms, B = 1 0.84 -> 0.73 ms.  At 64 tiles (the K >= 720 projections of the last blocks at 64
    // clips) narrower tiles measured SLOWER (1.60 -> 1.64 ms): the chunk period is the issue / commit chain, not the
    // MMA width, so more CTAs only add A re-reads and transform work.  AC_TC_NARROW=<tiles> moves the threshold, 0 = never.
    const TcWeight* pick(int M) const {
        if (tw.packed == nullptr) return nullptr;
        static const int below = [] { const char* e = getenv("AC_TC_NARROW"); return e ? atoi(e) : 64; }();
        const TcWeight* best = &tw;
        const int m_tiles = cdiv(M, 128);
        for (int l = 0; l < kNarrowLevels && m_tiles * best->n_tiles < below; ++l)
            if (narrow[l].packed != nullptr && narrow[l].n_tiles > best->n_tiles) best = &narrow[l];
        return best;
    }
};
struct BlockW {
    ConvBN expand, dw, project;
    float *se_wr, *se_br, *se_we, *se_be;
};

static int launch_se(const float* partial, int strips, float inv_hw, const BlockW& w, float* gate, int B, int C,
                     int nsq, cudaStream_t st) {
    static const int max_p = [] { const char* e = getenv("AC_SE_CLUSTER"); const int v = e ? atoi(e) : 8;
                                  return v >= 1 && v <= 8 ? v : 8; }();          // tuning aid
    int P = 1;
    while (P < max_p && C / (P * 2) >= 8) P *= 2;       // >= 2 channel quads per CTA
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * P); cfg.blockDim = dim3(kSeThreads);
    cfg.dynamicSmemBytes = (((C + nsq + 3) / 4) * 4 + kSeThreads * 4) * sizeof(float);
    cfg.stream = st;
    static const int solo_bytes = [] { const char* e = getenv("AC_SE_SOLO"); return e ? atoi(e) : 100000; }();   // bytes of one FC matrix; 0 = never
    if ((size_t)nsq * C * sizeof(float) <= (size_t)solo_bytes) {
        cudaLaunchAttribute at[1] = {pdl_attr()};
        cfg.attrs = at; cfg.numAttrs = 1;
        AC_TIMED("se", st);
        AC_CUDA(cudaLaunchKernelEx(&cfg, se_solo_kernel, partial, strips, inv_hw, (const float*)w.se_wr, (const float*)w.se_br,
                                   (const float*)w.se_we, (const float*)w.se_be, gate, C, nsq, P));
        AC_LAUNCHED("se_solo_kernel");
        return AC_OK;
    }
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = P; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1] = pdl_attr();
    cfg.attrs = attr; cfg.numAttrs = 2;
    AC_TIMED("se", st);
    AC_CUDA(cudaLaunchKernelEx(&cfg, se_kernel, partial, strips, inv_hw, (const float*)w.se_wr, (const float*)w.se_br,
                               (const float*)w.se_we, (const float*)w.se_be, gate, C, nsq));
    AC_LAUNCHED("se_kernel");
    return AC_OK;
}

static int out_size(int in, int k, int s, int lo, int hi) { return (in + lo + hi - k) / s + 1; }

}  // namespace ac

struct ac_effb2 {
    float* blob = nullptr;
    ac::ConvBN stem, head;
    std::vector<ac::BlockW> blocks;
};

namespace ac {

struct Dims { int H, W; };

// walks the plan and reports per-layer spatial sizes
static void walk(int n_mels, int n_frames, Dims& stem, std::vector<Dims>& in_dims, std::vector<Dims>& out_dims) {
    const Plan& P = plan();
    stem.H = out_size(n_mels, 3, 2, P.stem_pad_lo, P.stem_pad_hi);
    stem.W = out_size(n_frames, 3, 2, P.stem_pad_lo, P.stem_pad_hi);
    Dims d = stem;
    for (auto& b : P.blocks) {
        in_dims.push_back(d);
        d.H = out_size(d.H, b.k, b.s, b.pad_lo, b.pad_hi);
        d.W = out_size(d.W, b.k, b.s, b.pad_lo, b.pad_hi);
        out_dims.push_back(d);
    }
}

struct WsLayout { size_t x_elems, e_elems, d_elems, part_elems, gate_elems, head_elems; };

static WsLayout ws_layout(int batch, int n_mels, int n_frames) {
    const Plan& P = plan();
    Dims stem; std::vector<Dims> din, dout;
    walk(n_mels, n_frames, stem, din, dout);
    WsLayout L{};
    L.x_elems = (size_t)stem.H * stem.W * P.stem_out;
    for (size_t i = 0; i < P.blocks.size(); ++i) {
        auto& b = P.blocks[i];
        size_t pin = (size_t)din[i].H * din[i].W, pout = (size_t)dout[i].H * dout[i].W;
        L.x_elems = std::max(L.x_elems, pout * b.cout);
        if (b.expand != 1) L.e_elems = std::max(L.e_elems, pin * b.cexp());
        L.d_elems = std::max(L.d_elems, pout * b.cexp());
        size_t strips = (size_t)dwconv_tiles_per_clip(dout[i].H, dout[i].W, b.cexp(), b.k, b.s);
        L.part_elems = std::max(L.part_elems, strips * b.cexp());
        L.gate_elems = std::max(L.gate_elems, (size_t)b.cexp());
    }
    L.head_elems = (size_t)dout.back().H * dout.back().W * P.head_out;
    L.x_elems *= batch; L.e_elems *= batch; L.d_elems *= batch; L.part_elems *= batch;
    L.gate_elems *= batch; L.head_elems *= batch;
    return L;
}

}  // namespace ac

extern "C" {

int ac_effb2_num_tensors(void) {
    const ac::Plan& P = ac::plan();
    int n = 1 + 4;   // stem conv + bn
    for (auto& b : P.blocks) n += (b.expand != 1 ? 5 : 0) + 5 + 4 + 5;
    return n + 5;    // head conv + bn
}
int ac_effb2_out_dim(void) { return ac::plan().head_out; }
int ac_effb2_out_frames(int n_frames) {
    ac::Dims stem; std::vector<ac::Dims> a, b;
    ac::walk(64, n_frames, stem, a, b);
    return b.back().W;
}
int ac_effb2_block_info(int block, int* o) {
    const ac::Plan& P = ac::plan();
    if (block < 0 || block >= (int)P.blocks.size()) return (int)P.blocks.size();
    auto& b = P.blocks[block];
    o[0] = b.cin; o[1] = b.cout; o[2] = b.expand; o[3] = b.k; o[4] = b.s; o[5] = b.pad_lo; o[6] = b.pad_hi;
    o[7] = b.nsq; o[8] = b.skip;
    return (int)P.blocks.size();
}

size_t ac_effb2_workspace_bytes(int batch, int n_mels, int n_frames) {
    ac::WsLayout L = ac::ws_layout(batch, n_mels, n_frames);
    size_t fl = 2 * ac::align_up(L.x_elems, 64) + ac::align_up(L.e_elems, 64) + ac::align_up(L.d_elems, 64) +
                ac::align_up(L.part_elems, 64) + ac::align_up(L.gate_elems, 64) + ac::align_up(L.head_elems, 64);
    return fl * sizeof(float);
}

int ac_effb2_create(const float* const* t, const int64_t* numels, int n_tensors, void* stream, ac_effb2_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_effb2_create: null argument");
    AC_REQUIRE(n_tensors == ac_effb2_num_tensors(), "ac_effb2_create: expected %d tensors, got %d",
               ac_effb2_num_tensors(), n_tensors);
    const Plan& P = plan();
    cudaStream_t st = (cudaStream_t)stream;
    const float eps = 1e-3f;
    // ---- size the packed blob
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += align_up(n, 64); return o; };
    struct Off { size_t w, s, b, pk, npk[kNarrowLevels]; };
    auto take_cb = [&](size_t wn, size_t c) { Off o{}; o.w = take(wn); o.s = take(c); o.b = take(c); o.pk = 0; return o; };
    // 1x1 convolutions also get a tensor-core image (BN scale folded in, hi/lo split, swizzled; see gemm.cuh)
    auto take_pw = [&](int n, int k) {
        Off o = take_cb((size_t)n * k, n);
        o.pk = take(tc_packed_floats(n, k));
        for (int l = 0; l < kNarrowLevels; ++l) o.npk[l] = n > kNarrowCap[l] ? take(tc_packed_floats_bn(n, k, kNarrowCap[l])) : 0;
        return o;
    };
    Off stem_o = take_cb(9 * P.stem_out, P.stem_out);
    struct BOff { Off e, d, p; size_t wr, br, we, be; };
    std::vector<BOff> bo;
    for (auto& b : P.blocks) {
        BOff o{};
        int ce = b.cexp();
        if (b.expand != 1) o.e = take_pw(ce, b.cin);
        o.d = take_cb((size_t)b.k * b.k * ce, ce);
        o.wr = take((size_t)b.nsq * ce); o.br = take(b.nsq); o.we = take((size_t)ce * b.nsq); o.be = take(ce);
        o.p = take_pw(b.cout, ce);
        bo.push_back(o);
    }
    Off head_o = take_pw(P.head_out, P.head_in);
    ac_effb2_t* net = new ac_effb2_t();
    AC_CUDA(cudaMalloc(&net->blob, total * sizeof(float)));
    float* B0 = net->blob;
    int ti = 0;
    int rc = AC_OK;
    auto expect = [&](int64_t n, const char* what) {
        if (rc == AC_OK && numels[ti] != n) {
            set_error("ac_effb2_create: tensor %d (%s) has %lld elements, expected %lld", ti, what,
                      (long long)numels[ti], (long long)n);
            rc = AC_ERR_ARG;
        }
    };
    auto copy = [&](size_t off, int64_t n, const char* what) {
        expect(n, what);
        if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(B0 + off, t[ti], n * sizeof(float), cudaMemcpyDeviceToDevice, st), what);
        ++ti;
    };
    auto transposed = [&](size_t off, int rows, int cols, const char* what) {   // [rows][cols] -> [cols][rows]
        expect((int64_t)rows * cols, what);
        if (rc == AC_OK) {
            transpose_kernel<<<cdiv(rows * cols, 256), 256, 0, st>>>(t[ti], B0 + off, rows, cols);
            g_launches++;
        }
        ++ti;
    };
    auto bn = [&](const Off& o, int c, const char* what) {   // weight, bias, running_mean, running_var
        for (int q = 0; q < 4; ++q)
            if (rc == AC_OK && numels[ti + q] != c) {
                set_error("ac_effb2_create: tensor %d (%s) has %lld elements, expected %d", ti + q, what,
                          (long long)numels[ti + q], c);
                rc = AC_ERR_ARG;
            }
        if (rc == AC_OK) {
            bn_fold_kernel<<<cdiv(c, 256), 256, 0, st>>>(t[ti], t[ti + 1], t[ti + 2], t[ti + 3], eps, B0 + o.s, B0 + o.b, c);
            g_launches++;
        }
        ti += 4;
    };
    auto pack_pw = [&](const Off& o, int n, int k, ConvBN& cb) {
        cb = {B0 + o.w, B0 + o.s, B0 + o.b, TcWeight()};
        if (rc == AC_OK && k % 8 == 0) rc = tc_pack_weight(B0 + o.w, B0 + o.s, n, k, B0 + o.pk, st, &cb.tw);
        for (int l = 0; l < kNarrowLevels; ++l)
            if (rc == AC_OK && k % 8 == 0 && o.npk[l] != 0)
                rc = tc_pack_weight_bn(B0 + o.w, B0 + o.s, n, k, kNarrowCap[l], B0 + o.npk[l], st, &cb.narrow[l]);
    };
    transposed(stem_o.w, P.stem_out, 9, "_conv_stem.weight");
    bn(stem_o, P.stem_out, "_bn0");
    net->stem = {B0 + stem_o.w, B0 + stem_o.s, B0 + stem_o.b, TcWeight()};
    for (size_t i = 0; i < P.blocks.size(); ++i) {
        auto& b = P.blocks[i]; auto& o = bo[i];
        int ce = b.cexp();
        BlockW w{};
        if (b.expand != 1) {
            copy(o.e.w, (int64_t)ce * b.cin, "_expand_conv.weight");
            bn(o.e, ce, "_bn0");
            pack_pw(o.e, ce, b.cin, w.expand);
        }
        transposed(o.d.w, ce, b.k * b.k, "_depthwise_conv.weight");
        bn(o.d, ce, "_bn1");
        w.dw = {B0 + o.d.w, B0 + o.d.s, B0 + o.d.b, TcWeight()};
        copy(o.wr, (int64_t)b.nsq * ce, "_se_reduce.weight"); copy(o.br, b.nsq, "_se_reduce.bias");
        transposed(o.we, ce, b.nsq, "_se_expand.weight"); copy(o.be, ce, "_se_expand.bias");
        w.se_wr = B0 + o.wr; w.se_br = B0 + o.br; w.se_we = B0 + o.we; w.se_be = B0 + o.be;
        copy(o.p.w, (int64_t)b.cout * ce, "_project_conv.weight");
        bn(o.p, b.cout, "_bn2");
        pack_pw(o.p, b.cout, ce, w.project);
        net->blocks.push_back(w);
    }
    copy(head_o.w, (int64_t)P.head_out * P.head_in, "_conv_head.weight");
    bn(head_o, P.head_out, "_bn1");
    pack_pw(head_o, P.head_out, P.head_in, net->head);
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "ac_effb2_create pack kernels");
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_effb2_create sync");
    if (rc != AC_OK) { cudaFree(net->blob); delete net; return rc; }
    *out = net;
    return AC_OK;
}

void ac_effb2_destroy(ac_effb2_t* net) {
    if (!net) return;
    cudaFree(net->blob);
    delete net;
}

int ac_effb2_fwd(const ac_effb2_t* net, const float* lms, const float* gmax, float top_db, int B, int n_mels,
                 int n_frames, float* attn_emb, void* workspace, size_t ws_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 0 && B <= 65535, "ac_effb2_fwd: batch %d out of range", B);
    AC_REQUIRE(n_mels >= 8 && n_frames >= 32, "ac_effb2_fwd: input %dx%d too small", n_mels, n_frames);
    if (B == 0) return AC_OK;
    AC_REQUIRE(net && lms && attn_emb, "ac_effb2_fwd: null argument");
    AC_REQUIRE(workspace && ws_bytes >= ac_effb2_workspace_bytes(B, n_mels, n_frames),
               "ac_effb2_fwd: workspace too small (%zu < %zu)", ws_bytes, ac_effb2_workspace_bytes(B, n_mels, n_frames));
    const Plan& P = plan();
    cudaStream_t st = (cudaStream_t)stream;
    WsLayout L = ws_layout(B, n_mels, n_frames);
    float* p = (float*)workspace;
    float* X0 = p; p += align_up(L.x_elems, 64);
    float* X1 = p; p += align_up(L.x_elems, 64);
    float* E = p; p += align_up(L.e_elems, 64);
    float* D = p; p += align_up(L.d_elems, 64);
    float* PART = p; p += align_up(L.part_elems, 64);
    float* GATE = p; p += align_up(L.gate_elems, 64);
    float* HEAD = p; p += align_up(L.head_elems, 64);

    Dims stem; std::vector<Dims> din, dout;
    walk(n_mels, n_frames, stem, din, dout);
    // One MBConv block on clips [c0, c0 + nb) of the block's input `in` / output `out` (both [B, pixels, C]); the
    // expanded tensor E, the depthwise output D, the SE partial sums and gates are scratch reused by every chunk.
    auto run_block = [&](size_t i, const float* in, float* out, int nb) -> int {
        const BlockPlan& b = P.blocks[i];
        const BlockW& w = net->blocks[i];
        const int ce = b.cexp();
        const int pin = din[i].H * din[i].W, pout = dout[i].H * dout[i].W;
        const float* dw_in = in;
        if (b.expand != 1) {
            GemmArgs g; g.A = in; g.W = w.expand.w; g.C = E; g.M = nb * pin; g.N = ce; g.K = b.cin;
            g.cscale = w.expand.scale; g.cbias = w.expand.bias; g.act = ACT_SWISH;
            g.tw = w.expand.pick(g.M);
            int rc = gemm_tn(g, st); if (rc) return rc;
            dw_in = E;
        }
        DwArgs da;
        da.in = dw_in; da.out = D; da.partial = PART; da.w = w.dw.w; da.scale = w.dw.scale; da.bias = w.dw.bias;
        da.B = nb; da.Hi = din[i].H; da.Wi = din[i].W; da.Ho = dout[i].H; da.Wo = dout[i].W; da.C = ce;
        da.k = b.k; da.s = b.s; da.pad_lo = b.pad_lo;
        int rc = dwconv_tma(da, st); if (rc) return rc;
        const int strips = dwconv_tiles_per_clip(dout[i].H, dout[i].W, ce, b.k, b.s);
        rc = launch_se(PART, strips, 1.0f / (float)pout, w, GATE, nb, ce, b.nsq, st); if (rc) return rc;
        GemmArgs g; g.A = D; g.W = w.project.w; g.C = out; g.M = nb * pout; g.N = b.cout; g.K = ce;
        g.ascale = GATE; g.rows_per_group = pout; g.cscale = w.project.scale; g.cbias = w.project.bias;
        g.act = ACT_NONE; g.R = b.skip ? in : nullptr;
        g.tw = w.project.pick(g.M);
        return gemm_tn(g, st);
    };
    auto run_stem = [&](const float* lms_c, float* out, int nb) -> int {
        AC_REQUIRE(P.stem_out == 32, "effb2: the stem kernel is written for 32 output channels (got %d)", P.stem_out);
        const int tiles_w = cdiv(stem.W, kStemCols);
        AC_TIMED("stem", st);
        stem_kernel<<<(unsigned)(nb * stem.H * tiles_w), 256, 0, st>>>(lms_c, gmax, top_db, net->stem.w, net->stem.scale,
                                                                       net->stem.bias, out, n_mels, n_frames, stem.H, stem.W,
                                                                       P.stem_pad_lo, tiles_w);
        AC_LAUNCHED("stem_kernel");
        return AC_OK;
    };
    // ---- L2-resident head of the network (experiment).  The high-resolution blocks move 28 MB per clip between kernels (the 6x
    // expanded tensor is written by the expand GEMM, read and written by the depthwise kernel, read by the project GEMM);
    // at 64 clips no tensor survives in the 126 MB L2 from its producer to its consumer.  Running the stem and the first
    // `head_blocks` blocks one group of `chunk` clips at a time keeps a group's working set (<= 9.1 MB per clip) in L2, so
    // only the group's input and final output touch HBM.  Per-clip tensor sizes shrink along the chain, hence group g's
    // intermediates (at clip offset g * chunk in the ping-pong buffers) never reach the finished outputs of groups < g
    // nor the unread stem input of groups > g.
    // MEASURED (scripts/effb2_chunk_sweep.py, 64 clips): slower, not faster -- 5 blocks x 8 clips: encoder 2.75 -> 3.36 ms,
    // sum of kernel times 3.17 -> 4.44 ms.  Every extra launch of the tensor-core GEMM / SE kernels costs ~8 us of fixed
    // prologue + drain, which outweighs whatever DRAM traffic the L2 residency saves.  Off by default; the switch stays
    // for the record: AC_EFFB2_CHUNK="blocks,clips".
    int head_blocks = 0, chunk = 0;
    if (const char* e = getenv("AC_EFFB2_CHUNK")) sscanf(e, "%d,%d", &head_blocks, &chunk);
    head_blocks = std::max(0, std::min(head_blocks, (int)P.blocks.size()));
    if (chunk <= 0 || chunk >= B) head_blocks = 0;
    float* cur = X0; float* nxt = X1;
    if (head_blocks == 0) {
        int rc = run_stem(lms, X0, B); if (rc) return rc;
    } else {
        const size_t s_stem = (size_t)stem.H * stem.W * P.stem_out;
        for (int c0 = 0; c0 < B; c0 += chunk) {
            const int nb = std::min(chunk, B - c0);
            float* a = X0; float* bb = X1;
            int rc = run_stem(lms + (size_t)c0 * n_mels * n_frames, a + (size_t)c0 * s_stem, nb); if (rc) return rc;
            size_t s_in = s_stem;
            for (int i = 0; i < head_blocks; ++i) {
                const size_t s_out = (size_t)dout[i].H * dout[i].W * P.blocks[i].cout;
                rc = run_block(i, a + (size_t)c0 * s_in, bb + (size_t)c0 * s_out, nb); if (rc) return rc;
                std::swap(a, bb);
                s_in = s_out;
            }
        }
        if (head_blocks & 1) std::swap(cur, nxt);
    }
    for (size_t i = head_blocks; i < P.blocks.size(); ++i) {
        int rc = run_block(i, cur, nxt, B); if (rc) return rc;
        std::swap(cur, nxt);
    }
    const Dims last = dout.back();
    {
        GemmArgs g; g.A = cur; g.W = net->head.w; g.C = HEAD; g.M = B * last.H * last.W; g.N = P.head_out;
        g.K = P.head_in; g.cscale = net->head.scale; g.cbias = net->head.bias; g.act = ACT_SWISH;
        g.tw = net->head.pick(g.M);
        int rc = gemm_tn(g, st); if (rc) return rc;
        int64_t total = (int64_t)B * last.W * P.head_out / 4;
        AC_TIMED("freq_mean", st);
        freq_mean_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(HEAD, attn_emb, last.H, last.W, P.head_out, total);
        AC_LAUNCHED("freq_mean_kernel");
    }
    return AC_OK;
}

int ac_masked_mean(const float* x, const int64_t* lens, int batch, int T, int D, float* out, void* stream) {
    if (batch == 0 || D == 0) return AC_OK;
    AC_REQUIRE(x && lens && out, "ac_masked_mean: null argument");
    dim3 grid(ac::cdiv(D, 128), batch);
    ac::masked_mean_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, lens, T, D, out);
    AC_LAUNCHED("masked_mean_kernel");
    return AC_OK;
}

}  // extern "C"
