// Cnn14 (PANNs) audio encoder body for sm_100a.
//
// Replaces captioning/models/cnn_encoder.py:414-464 `Cnn14Encoder.forward` after the log-mel front-end (HF copy:
// hf_wrapper.py:1259-1304), eval mode: bn0 over the mel axis -> 6 x ConvBlock (cnn_encoder.py:32-75:
// [conv3x3 -> BatchNorm -> ReLU] x 2 -> avg_pool 2x2, block 6 without pooling) -> mean over mel -> attn_emb;
// fc_emb = relu(fc1(max_with_lens + mean_with_lens)) (model_util.py:41-84).  Dropout is the identity in eval mode.
//
// Layout: activations NHWC fp32 [B, H = time, W = mel, C].  The 11 convolutions with Cin >= 64 (99.8 % of the
// 40 GFLOP per clip) run as implicit GEMMs on the tcgen05 pipeline of gemm_tc.cu (`conv3x3_tc`: 4-D TMA boxes
// shifted per tap, 3xTF32 split, BN scale folded into the packed weight, bias + ReLU in the epilogue).  The first
// convolution (Cin = 1, 9 MACs per output) is an HBM-write-bound SIMT kernel that also applies bn0 and the
// [B, mel, T] -> [B, T, mel] transposition while loading.
#include <algorithm>
#include <vector>

#include <cuda_bf16.h>

#include "gemm.cuh"
#include "train_ops.cuh"

namespace ac {

constexpr int kCnn14Blocks = 6;
constexpr int kCnn14Ch[kCnn14Blocks + 1] = {1, 64, 128, 256, 512, 1024, 2048};
constexpr float kCnn14BnEps = 1e-5f;   // nn.BatchNorm2d default
constexpr int kC1Time = 16;            // time steps per CTA of the first convolution
constexpr int kC1Threads = 256;

__global__ void cnn14_bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                     float* __restrict__ scale, float* __restrict__ bias, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float s = gamma[i] / sqrtf(var[i] + eps);
        scale[i] = s;
        bias[i] = beta[i] - mean[i] * s;
    }
}
// conv_block1.conv1.weight [64, 1, 3, 3] -> [9][64]
__global__ void cnn14_w1_kernel(const float* __restrict__ w, float* __restrict__ o, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * C) o[(i % 9) * C + i / 9] = w[i];
}

// First convolution: lms [B, M = mel, T] (dB) -> out [B, T, M, 64] = relu(bn1(conv3x3(bn0(lms^T)))).
// A CTA owns kC1Time time steps of one clip: the (kC1Time + 2) x (M + 2) input patch (bn0 applied, zero padding
// outside the image) sits in shared memory; 16 threads serve one pixel (4 output channels each, 36 weights in
// registers), so a pixel's 64 channels leave as one 256-byte run.
// (OT = float, or __nv_bfloat16 in the bf16 precision mode: the same arithmetic, the store rounds)
template <typename OT>
__global__ void __launch_bounds__(kC1Threads)
cnn14_conv1_kernel(const float* __restrict__ lms, const float* __restrict__ s0, const float* __restrict__ t0,
                   const float* __restrict__ w /*[9][64]*/, const float* __restrict__ scale,
                   const float* __restrict__ bias, OT* __restrict__ out, int M, int T) {
    extern __shared__ float patch[];                     // [(kC1Time + 2)][M + 2]
    const int b = blockIdx.y, tb = blockIdx.x * kC1Time;
    const int PW = M + 2;
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < (kC1Time + 2) * PW; i += kC1Threads) {
        // i -> (m, dt) with dt fastest so that consecutive threads read consecutive time steps of one mel row
        const int dt = i % (kC1Time + 2), mm = i / (kC1Time + 2);
        const int t = tb + dt - 1, m = mm - 1;
        float v = 0.0f;
        if (t >= 0 && t < T && m >= 0 && m < M) v = __ldg(lms + ((size_t)b * M + m) * T + t) * __ldg(s0 + m) + __ldg(t0 + m);
        patch[dt * PW + mm] = v;
    }
    const int cg = threadIdx.x & 15, slot = threadIdx.x >> 4;   // 4 channels, pixel slot 0..15
    float wr[9][4], sc[4], bi[4];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w + k * 64) + cg);
        wr[k][0] = v.x; wr[k][1] = v.y; wr[k][2] = v.z; wr[k][3] = v.w;
    }
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(scale) + cg), c = __ldg(reinterpret_cast<const float4*>(bias) + cg);
        sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; bi[0] = c.x; bi[1] = c.y; bi[2] = c.z; bi[3] = c.w;
    }
    __syncthreads();
    const int npix = kC1Time * M;
    for (int px = slot; px < npix; px += kC1Threads / 16) {
        const int dt = px / M, m = px % M;
        if (tb + dt >= T) break;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float x = patch[(dt + ky) * PW + m + kx];
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] = fmaf(x, wr[ky * 3 + kx][e], acc[e]);
            }
        float4 o;
        o.x = fmaxf(fmaf(acc[0], sc[0], bi[0]), 0.f); o.y = fmaxf(fmaf(acc[1], sc[1], bi[1]), 0.f);
        o.z = fmaxf(fmaf(acc[2], sc[2], bi[2]), 0.f); o.w = fmaxf(fmaf(acc[3], sc[3], bi[3]), 0.f);
        OT* dst = out + (((size_t)b * T + tb + dt) * M + m) * 64 + 4 * cg;
        if constexpr (sizeof(OT) == 4) {
            *reinterpret_cast<float4*>(dst) = o;
        } else {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
}

// ---- bf16 precision mode: the pooling / dropout / tail kernels on bf16 NHWC activations (8 channels = 16 bytes per thread)
__device__ __forceinline__ void bf16x8_to_float(const uint4 v, float (&f)[8]) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[e]));
        f[2 * e] = p.x; f[2 * e + 1] = p.y;
    }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&f)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        u[e] = *reinterpret_cast<const uint32_t*>(&p);
    }
    return make_uint4(u[0], u[1], u[2], u[3]);
}
// avg_pool2d 2x2 (+ the ConvBlock dropout in training): same element indices for the dropout stream as the fp32 kernel
__global__ void __launch_bounds__(256)
cnn14_avgpool_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int Ho, int Wo,
                          int64_t total, Dropout dp, uint32_t site) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C8);
    int64_t r = i / C8;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int64_t b = r / Ho;
    const uint4* p = in + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C8 + c;
    float a[8], bb[8], cc[8], d[8], o[8];
    bf16x8_to_float(__ldg(p), a); bf16x8_to_float(__ldg(p + C8), bb);
    bf16x8_to_float(__ldg(p + (size_t)W * C8), cc); bf16x8_to_float(__ldg(p + (size_t)W * C8 + C8), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        o[e] = 0.25f * (a[e] + bb[e] + cc[e] + d[e]);
        if (dp.p > 0.0f) o[e] *= drop_scale(dp.seed, site, (uint64_t)i * 8 + e, dp.p);
    }
    out[i] = float_to_bf16x8(o);
}
__global__ void __launch_bounds__(256)
cnn14_dropout_bf16_kernel(uint4* __restrict__ x, int64_t total, Dropout dp, uint32_t site) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float f[8];
    bf16x8_to_float(x[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= drop_scale(dp.seed, site, (uint64_t)i * 8 + e, dp.p);
    x[i] = float_to_bf16x8(f);
}

// avg_pool2d(kernel 2x2, stride 2, floor): in [B, H, W, C] -> out [B, H/2, W/2, C]; one float4 per thread.
__global__ void __launch_bounds__(256)
cnn14_avgpool_kernel(const float4* __restrict__ in, float4* __restrict__ out, int H, int W, int C4, int Ho, int Wo,
                     int64_t total, Dropout dp, uint32_t site) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int64_t b = r / Ho;
    const float4* p = in + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 a = __ldg(p), bb = __ldg(p + C4), cc = __ldg(p + (size_t)W * C4), d = __ldg(p + (size_t)W * C4 + C4);
    float4 o = make_float4(0.25f * (a.x + bb.x + cc.x + d.x), 0.25f * (a.y + bb.y + cc.y + d.y),
                           0.25f * (a.z + bb.z + cc.z + d.z), 0.25f * (a.w + bb.w + cc.w + d.w));
    if (dp.p > 0.0f) {     // F.dropout(x, p=0.2, training=self.training) after the block (cnn_encoder.py:432-456)
        o.x *= drop_scale(dp.seed, site, (uint64_t)i * 4, dp.p); o.y *= drop_scale(dp.seed, site, (uint64_t)i * 4 + 1, dp.p);
        o.z *= drop_scale(dp.seed, site, (uint64_t)i * 4 + 2, dp.p); o.w *= drop_scale(dp.seed, site, (uint64_t)i * 4 + 3, dp.p);
    }
    out[i] = o;
}

// bf16 forms of the SED pooling (avg + max over ph x pw) and of the mean over mel (-> fp32 for fc1)
__global__ void __launch_bounds__(256)
cnn_avgmax_pool_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int Ho, int Wo, int ph,
                            int pw, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C8);
    int64_t r = i / C8;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int64_t b = r / Ho;
    float sum[8], mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sum[e] = 0.f; mx[e] = -INFINITY; }
    for (int dy = 0; dy < ph; ++dy)
        for (int dx = 0; dx < pw; ++dx) {
            float v[8];
            bf16x8_to_float(__ldg(in + (((size_t)b * H + ho * ph + dy) * W + wo * pw + dx) * C8 + c), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) { sum[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
        }
    const float inv = 1.0f / (float)(ph * pw);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = sum[e] * inv + mx[e];
    out[i] = float_to_bf16x8(o);
}
__global__ void __launch_bounds__(256)
cnn_wmean_bf16_kernel(const uint4* __restrict__ y, float4* __restrict__ out, int W, int C8, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C8);
    const int64_t bh = i / C8;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int w = 0; w < W; ++w) {
        float v[8];
        bf16x8_to_float(__ldg(y + ((size_t)bh * W + w) * C8 + c), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += v[e];
    }
    const float inv = 1.0f / (float)W;
    out[2 * i] = make_float4(s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv);
    out[2 * i + 1] = make_float4(s[4] * inv, s[5] * inv, s[6] * inv, s[7] * inv);
}

// ConvBlock pooling of the SED tagger (hf_wrapper.py:1204-1212, pool_type 'avg+max'): avg_pool2d + max_pool2d over
// (ph x pw) windows, stride = window, floor.  in [B, H, W, C] -> out [B, H/ph, W/pw, C]; one float4 per thread.
__global__ void __launch_bounds__(256)
cnn_avgmax_pool_kernel(const float4* __restrict__ in, float4* __restrict__ out, int H, int W, int C4, int Ho, int Wo, int ph,
                       int pw, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int64_t b = r / Ho;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < ph; ++dy)
        for (int dx = 0; dx < pw; ++dx) {
            const float4 v = __ldg(in + (((size_t)b * H + ho * ph + dy) * W + wo * pw + dx) * C4 + c);
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
        }
    const float inv = 1.0f / (float)(ph * pw);
    out[i] = make_float4(sum.x * inv + mx.x, sum.y * inv + mx.y, sum.z * inv + mx.z, sum.w * inv + mx.w);
}

// y [B, H, W, C] -> out [B, H, C] = mean over W   (torch.mean(x, dim=3) then transpose(1, 2))
__global__ void __launch_bounds__(256)
cnn_wmean_kernel(const float4* __restrict__ y, float4* __restrict__ out, int W, int C4, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C4);
    const int64_t bh = i / C4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < W; ++w) {
        const float4 v = __ldg(y + ((size_t)bh * W + w) * C4 + c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.0f / (float)W;
    out[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

// logits [rows, ld] (first `classes` columns valid) -> prob = clamp(sigmoid(logit), 1e-7, 1) in place
__global__ void sed_sigmoid_kernel(float* __restrict__ x, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) x[i] = fminf(fmaxf(1.0f / (1.0f + expf(-x[i])), 1e-7f), 1.0f);
}

__global__ void sed_fill_kernel(int64_t* __restrict__ x, int64_t v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// Double threshold (hysteresis) along time for every (clip, class) column: a maximal run of prob > low is kept when
// it contains a value > high (hf_wrapper.py:123-162 `double_threshold`; at the repeat-upsampled frame resolution the
// n_connect = 1 merge can never fire because runs are >= 4 frames apart).  prob [B, S, ld] -> labels [B, S, classes] u8.
// Every kept run is also appended to `runs` as (clip, class, first segment, one past the last segment) through an atomic
// cursor (order unspecified; the pairwise rule that consumes them is order-free), so the host needs a few hundred bytes
// instead of the label matrix.
__global__ void sed_hysteresis_kernel(const float* __restrict__ prob, unsigned char* __restrict__ labels, int S, int ld,
                                      int classes, float high, float low, int total, int4* __restrict__ runs, int max_runs,
                                      int* __restrict__ n_runs) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = i % classes, b = i / classes;
    const float* p = prob + (size_t)b * S * ld + c;
    unsigned char* l = labels + (size_t)b * S * classes + c;
    int start = -1; bool has_high = false;
    for (int s = 0; s <= S; ++s) {
        const float v = s < S ? p[(size_t)s * ld] : -1.0f;
        if (v > low) {
            if (start < 0) { start = s; has_high = false; }
            has_high = has_high || v > high;
        } else if (start >= 0) {
            const unsigned char keep = has_high ? 1 : 0;
            for (int q = start; q < s; ++q) l[(size_t)q * classes] = keep;
            if (keep && n_runs != nullptr) {
                const int idx = atomicAdd(n_runs, 1);
                if (idx < max_runs) runs[idx] = make_int4(b, c, start, s);
            }
            start = -1;
        }
        if (s < S && !(v > low)) l[(size_t)s * classes] = 0;
    }
}

// y [B, H, W, C] -> attn_emb [B, H, C] = mean over W (torch.mean(x, dim=3), 'b c t f -> b t c') and
// pooled [B, C] = max_{t < len} attn_emb + (sum_{t < len} attn_emb) / len   (max_with_lens + mean_with_lens)
template <typename IT>
__global__ void __launch_bounds__(128)
cnn14_tail_kernel(const IT* __restrict__ y, const int64_t* __restrict__ lens, float* __restrict__ attn,
                  float* __restrict__ pooled, int H, int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t len = lens[b];
    const float inv_w = 1.0f / (float)W;
    float mx = -INFINITY, sum = 0.0f;
    for (int t = 0; t < H; ++t) {
        float s = 0.0f;
        for (int w = 0; w < W; ++w) {
            if constexpr (sizeof(IT) == 4) s += __ldg(y + (((size_t)b * H + t) * W + w) * C + c);
            else s += __bfloat162float(y[(((size_t)b * H + t) * W + w) * C + c]);
        }
        s *= inv_w;
        attn[((size_t)b * H + t) * C + c] = s;
        if (t < len) { mx = fmaxf(mx, s); sum += s; }
    }
    pooled[(size_t)b * C + c] = mx + sum / (float)len;
}

struct Cnn14Conv { float* scale; float* bias; TcWeight tw; ConvBf16Weight bw; int cin, cout; };

}  // namespace ac

struct ac_cnn14 {
    float* blob = nullptr;
    void* blob_bf16 = nullptr;    // bf16 images of the convolution weights (precision mode 16)
    float *bn0_s, *bn0_b, *w1, *s1, *b1;
    ac::Cnn14Conv conv[2 * ac::kCnn14Blocks - 1];   // block1.conv2, block2.conv1, ... block6.conv2
    float *fc_w, *fc_b;
    ac::TcWeight fc_tw;
    int conv_passes = 3;          // 3 = 3xTF32 (fp32-level), 1 = plain TF32, 16 = bf16 activations + weights (ac_cnn14_set_precision)
    int sm_limit = 0;             // persistent CTAs of the bf16 convolutions (0 = one per SM): ac_cnn14_set_sm_limit
};

namespace ac {
struct Cnn14Dims { int H, W; };
static void cnn14_walk(int n_mels, int n_frames, Cnn14Dims (&d)[kCnn14Blocks]) {   // input dims of each block
    int H = n_frames, W = n_mels;
    for (int i = 0; i < kCnn14Blocks; ++i) { d[i] = {H, W}; if (i + 1 < kCnn14Blocks) { H /= 2; W /= 2; } }
}
static size_t cnn14_act_elems(int batch, int n_mels, int n_frames) {
    Cnn14Dims d[kCnn14Blocks];
    cnn14_walk(n_mels, n_frames, d);
    size_t m = 0;
    for (int i = 0; i < kCnn14Blocks; ++i) m = std::max(m, (size_t)d[i].H * d[i].W * kCnn14Ch[i + 1]);
    return align_up(m * batch, 64);
}
template <typename... Args>
static int launch_pdl(void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1] = {pdl_attr()};
    cfg.attrs = at; cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, kernel, args...), "cudaLaunchKernelEx");
}
}  // namespace ac

extern "C" {

int ac_cnn14_num_tensors(void) { return 4 + ac::kCnn14Blocks * 10 + 2; }
int ac_cnn14_out_dim(void) { return ac::kCnn14Ch[ac::kCnn14Blocks]; }
int ac_cnn14_out_frames(int n_frames) {
    for (int i = 0; i + 1 < ac::kCnn14Blocks; ++i) n_frames /= 2;
    return n_frames;
}
size_t ac_cnn14_workspace_bytes(int batch, int n_mels, int n_frames) {
    return (2 * ac::cnn14_act_elems(batch, n_mels, n_frames) + ac::align_up((size_t)batch * ac_cnn14_out_dim(), 64)) * sizeof(float);
}

int ac_cnn14_create(const float* const* t, const int64_t* numels, int n_tensors, void* stream, ac_cnn14_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_cnn14_create: null argument");
    AC_REQUIRE(n_tensors == ac_cnn14_num_tensors(), "ac_cnn14_create: expected %d tensors, got %d", ac_cnn14_num_tensors(),
               n_tensors);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = ac_cnn14_out_dim();
    // expected element counts, in the order documented in the header
    std::vector<int64_t> want;
    for (int q = 0; q < 4; ++q) want.push_back(64);
    for (int i = 0; i < kCnn14Blocks; ++i) {
        const int ci = kCnn14Ch[i], co = kCnn14Ch[i + 1];
        want.push_back((int64_t)co * ci * 9); want.push_back((int64_t)co * co * 9);
        for (int q = 0; q < 8; ++q) want.push_back(co);
    }
    want.push_back((int64_t)D * D); want.push_back(D);
    for (int i = 0; i < n_tensors; ++i)
        AC_REQUIRE(numels[i] == want[i], "ac_cnn14_create: tensor %d has %lld elements, expected %lld", i,
                   (long long)numels[i], (long long)want[i]);

    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += align_up(n, 64); return o; };
    const size_t o_bn0s = take(64), o_bn0b = take(64), o_w1 = take(9 * 64), o_s1 = take(64), o_b1 = take(64);
    struct Off { size_t s, b, pk; };
    Off co_[2 * kCnn14Blocks - 1];
    int cin_[2 * kCnn14Blocks - 1], cout_[2 * kCnn14Blocks - 1];
    size_t perm_max = 0;
    {
        int l = 0;
        for (int i = 0; i < kCnn14Blocks; ++i)
            for (int j = 0; j < 2; ++j) {
                if (i == 0 && j == 0) continue;
                cin_[l] = j == 0 ? kCnn14Ch[i] : kCnn14Ch[i + 1]; cout_[l] = kCnn14Ch[i + 1];
                co_[l].s = take(cout_[l]); co_[l].b = take(cout_[l]);
                co_[l].pk = take(tc_packed_floats(cout_[l], 9 * cin_[l]));
                perm_max = std::max(perm_max, (size_t)cout_[l] * 9 * cin_[l]);
                ++l;
            }
    }
    const size_t o_fcw = take((size_t)D * D), o_fcb = take(D), o_fcpk = take(tc_packed_floats(D, D));
    ac_cnn14_t* net = new ac_cnn14_t();
    float* perm = nullptr;
    size_t bf_total = 0, bf_off[2 * kCnn14Blocks - 1];
    for (int q = 0; q < 2 * kCnn14Blocks - 1; ++q) { bf_off[q] = bf_total; bf_total += align_up(conv_bf16_packed_elems(cout_[q], cin_[q]), 64); }
    int rc = check_cuda(cudaMalloc(&net->blob, total * sizeof(float)), "ac_cnn14_create: cudaMalloc(weights)");
    if (rc == AC_OK) rc = check_cuda(cudaMalloc(&net->blob_bf16, bf_total * 2), "ac_cnn14_create: cudaMalloc(bf16 weights)");
    if (rc == AC_OK) rc = check_cuda(cudaMalloc(&perm, perm_max * sizeof(float)), "ac_cnn14_create: cudaMalloc(scratch)");
    if (rc != AC_OK) { cudaFree(net->blob); cudaFree(net->blob_bf16); delete net; return rc; }
    float* B0 = net->blob;
    auto fold = [&](int ti, int c, float* s, float* b) {
        cnn14_bn_fold_kernel<<<cdiv(c, 256), 256, 0, st>>>(t[ti], t[ti + 1], t[ti + 2], t[ti + 3], kCnn14BnEps, s, b, c);
        g_launches++;
    };
    net->bn0_s = B0 + o_bn0s; net->bn0_b = B0 + o_bn0b; net->w1 = B0 + o_w1; net->s1 = B0 + o_s1; net->b1 = B0 + o_b1;
    fold(0, 64, net->bn0_s, net->bn0_b);
    int l = 0;
    for (int i = 0; i < kCnn14Blocks && rc == AC_OK; ++i) {
        const int base = 4 + i * 10;   // conv1.weight, conv2.weight, bn1 x4, bn2 x4
        for (int j = 0; j < 2 && rc == AC_OK; ++j) {
            const int bn_ti = base + 2 + 4 * j;
            if (i == 0 && j == 0) {
                cnn14_w1_kernel<<<cdiv(9 * 64, 256), 256, 0, st>>>(t[base], net->w1, 64);
                g_launches++;
                fold(bn_ti, 64, net->s1, net->b1);
                continue;
            }
            Cnn14Conv& c = net->conv[l];
            c.cin = cin_[l]; c.cout = cout_[l]; c.scale = B0 + co_[l].s; c.bias = B0 + co_[l].b;
            fold(bn_ti, c.cout, c.scale, c.bias);
            rc = conv3x3_permute_weight(t[base + j], perm, c.cout, c.cin, st);
            if (rc == AC_OK) rc = tc_pack_weight(perm, c.scale, c.cout, 9 * c.cin, B0 + co_[l].pk, st, &c.tw);
            if (rc == AC_OK) rc = conv_bf16_pack(perm, c.scale, c.cout, c.cin, (uint16_t*)net->blob_bf16 + bf_off[l], st, &c.bw);
            ++l;
        }
    }
    const int fc_ti = 4 + kCnn14Blocks * 10;
    net->fc_w = B0 + o_fcw; net->fc_b = B0 + o_fcb;
    if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(net->fc_w, t[fc_ti], (size_t)D * D * sizeof(float), cudaMemcpyDeviceToDevice, st), "fc1.weight");
    if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(net->fc_b, t[fc_ti + 1], (size_t)D * sizeof(float), cudaMemcpyDeviceToDevice, st), "fc1.bias");
    if (rc == AC_OK) rc = tc_pack_weight(net->fc_w, nullptr, D, D, B0 + o_fcpk, st, &net->fc_tw);
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "ac_cnn14_create pack kernels");
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_cnn14_create sync");
    cudaFree(perm);
    if (rc != AC_OK) { cudaFree(net->blob); cudaFree(net->blob_bf16); delete net; return rc; }
    *out = net;
    return AC_OK;
}

void ac_cnn14_destroy(ac_cnn14_t* net) {
    if (!net) return;
    cudaFree(net->blob);
    cudaFree(net->blob_bf16);
    delete net;
}

// tf32_passes = 3 (default): every 3x3 convolution product as three TF32 MMAs -- fp32-level accuracy, the mode all
// fp32 parity numbers are quoted in; 1: plain TF32 operands (10-bit mantissa, fp32 accumulate), the tensor-core mode for
// the configurations BASELINE.json states in bf16 (training, temporal captioner).
int ac_cnn14_set_precision(ac_cnn14_t* net, int tf32_passes) {
    using namespace ac;
    AC_REQUIRE(net && (tf32_passes == 1 || tf32_passes == 3 || tf32_passes == 16),
               "ac_cnn14_set_precision: mode must be 3 (3xTF32), 1 (TF32) or 16 (bf16)");
    net->conv_passes = tf32_passes;
    return AC_OK;
}

// Caps the persistent grid of the bf16 convolutions at n CTAs (0 = one per SM).  The training step runs the frozen
// encoder of batch i+1 on a second stream next to the (latency-bound, few-CTA) trainable part of batch i
// (audiocaption_b200/train_step.py `TrainStep.prefetch`): leaving some SMs to that chain keeps it moving while the
// convolutions hold the rest.
int ac_cnn14_set_sm_limit(ac_cnn14_t* net, int n) {
    using namespace ac;
    AC_REQUIRE(net && n >= 0, "ac_cnn14_set_sm_limit: bad argument");
    net->sm_limit = n;
    return AC_OK;
}

int ac_cnn14_fwd(const ac_cnn14_t* net, const float* lms, int B, int n_mels, int n_frames, const int64_t* lens,
                 float* attn_emb, float* fc_emb, void* workspace, size_t ws_bytes, void* stream) {
    return ac_cnn14_fwd_train(net, lms, B, n_mels, n_frames, lens, 0.0f, 0.0f, 0, attn_emb, fc_emb, workspace, ws_bytes, stream);
}

// Train-mode forward of the FROZEN encoder (eg_configs/*/waveform/cnn14rnn_trm.yaml: freeze_cnn + freeze_cnn_bn): BatchNorm
// stays in eval mode (folded), but the functional dropouts of cnn_encoder.py:432-456 are active -- p_conv after every
// ConvBlock (0.2 in the reference), p_fc around fc1 (0.5).  Masks come from the counter RNG (seed); p = 0 is ac_cnn14_fwd.
int ac_cnn14_fwd_train(const ac_cnn14_t* net, const float* lms, int B, int n_mels, int n_frames, const int64_t* lens,
                       float p_conv, float p_fc, uint64_t seed, float* attn_emb, float* fc_emb, void* workspace,
                       size_t ws_bytes, void* stream) {
    using namespace ac;
    const Dropout dpc{p_conv, seed}, dpf{p_fc, seed};
    constexpr uint32_t kSiteCnn = 96;      // dropout streams 96..101: after block 1..6; 102, 103: around fc1
    AC_REQUIRE(B >= 0 && B <= 65535, "ac_cnn14_fwd: batch %d out of range", B);
    AC_REQUIRE(n_mels == 64, "ac_cnn14_fwd: bn0 is defined over 64 mel bins, got %d", n_mels);
    AC_REQUIRE(n_frames >= 32, "ac_cnn14_fwd: %d frames is fewer than the down-sampling ratio 32", n_frames);
    if (B == 0) return AC_OK;
    AC_REQUIRE(net && lms && lens && attn_emb && fc_emb, "ac_cnn14_fwd: null argument");
    AC_REQUIRE(workspace && ws_bytes >= ac_cnn14_workspace_bytes(B, n_mels, n_frames),
               "ac_cnn14_fwd: workspace too small (%zu < %zu)", ws_bytes, ac_cnn14_workspace_bytes(B, n_mels, n_frames));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t act = cnn14_act_elems(B, n_mels, n_frames);
    float* cur = (float*)workspace;
    float* nxt = cur + act;
    float* pooled = nxt + act;
    Cnn14Dims d[kCnn14Blocks];
    cnn14_walk(n_mels, n_frames, d);
    int rc;
    if (net->conv_passes == 16) {
        // ---- bf16 precision mode: every activation between bn0 and the pooled output lives in bf16 NHWC (the two fp32
        // activation buffers of the workspace hold them), the convolutions run on conv3x3_bf16 (bf16 x bf16 -> fp32 in TMEM)
        __nv_bfloat16* bcur = reinterpret_cast<__nv_bfloat16*>(cur);
        __nv_bfloat16* bnxt = reinterpret_cast<__nv_bfloat16*>(nxt);
        {
            AC_TIMED("cnn14_conv1", st);
            dim3 grid(cdiv(n_frames, kC1Time), B);
            const size_t smem = (size_t)(kC1Time + 2) * (n_mels + 2) * sizeof(float);
            rc = launch_pdl(cnn14_conv1_kernel<__nv_bfloat16>, grid, dim3(kC1Threads), smem, st, lms, (const float*)net->bn0_s,
                            (const float*)net->bn0_b, (const float*)net->w1, (const float*)net->s1, (const float*)net->b1, bcur,
                            n_mels, n_frames);
            if (rc) return rc;
            AC_LAUNCHED("cnn14_conv1_kernel");
        }
        int l = 0;
        for (int i = 0; i < kCnn14Blocks; ++i) {
            for (int j = 0; j < 2; ++j) {
                if (i == 0 && j == 0) continue;
                const Cnn14Conv& c = net->conv[l++];
                ConvBf16Args a; a.in = bcur; a.out = bnxt; a.B = B; a.H = d[i].H; a.W = d[i].W; a.Cin = c.cin; a.Cout = c.cout;
                a.bias = c.bias; a.w = &c.bw; a.act = ACT_RELU; a.max_ctas = net->sm_limit;
                rc = conv3x3_bf16(a, st); if (rc) return rc;
                std::swap(bcur, bnxt);
            }
            if (i + 1 < kCnn14Blocks) {
                const int C8 = kCnn14Ch[i + 1] / 8, Ho = d[i].H / 2, Wo = d[i].W / 2;
                const int64_t total = (int64_t)B * Ho * Wo * C8;
                AC_TIMED("cnn14_avgpool", st);
                rc = launch_pdl(cnn14_avgpool_bf16_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st,
                                (const uint4*)bcur, (uint4*)bnxt, d[i].H, d[i].W, C8, Ho, Wo, total, dpc, kSiteCnn + (uint32_t)i);
                if (rc) return rc;
                AC_LAUNCHED("cnn14_avgpool_kernel");
                std::swap(bcur, bnxt);
            }
        }
        const Cnn14Dims last = d[kCnn14Blocks - 1];
        const int D = ac_cnn14_out_dim();
        if (dpc.p > 0.0f) {
            const int64_t total = (int64_t)B * last.H * last.W * D / 8;
            rc = launch_pdl(cnn14_dropout_bf16_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, (uint4*)bcur, total, dpc,
                            kSiteCnn + (uint32_t)kCnn14Blocks - 1);
            if (rc) return rc;
            AC_LAUNCHED("cnn14_dropout_kernel");
        }
        {
            AC_TIMED("cnn14_tail", st);
            rc = launch_pdl(cnn14_tail_kernel<__nv_bfloat16>, dim3(cdiv(D, 128), B), dim3(128), 0, st, (const __nv_bfloat16*)bcur, lens,
                            attn_emb, pooled, last.H, last.W, D);
            if (rc) return rc;
            AC_LAUNCHED("cnn14_tail_kernel");
        }
        rc = dropout_apply(pooled, 0, (int64_t)B * D, dpf, kSiteCnn + kCnn14Blocks, st); if (rc) return rc;
        GemmArgs g; g.A = pooled; g.W = net->fc_w; g.C = fc_emb; g.M = B; g.N = D; g.K = D; g.cbias = net->fc_b; g.act = ACT_RELU;
        g.tw = &net->fc_tw;
        rc = gemm_tn(g, st); if (rc) return rc;
        return dropout_apply(fc_emb, 0, (int64_t)B * D, dpf, kSiteCnn + kCnn14Blocks + 1, st);
    }
    {
        AC_TIMED("cnn14_conv1", st);
        dim3 grid(cdiv(n_frames, kC1Time), B);
        const size_t smem = (size_t)(kC1Time + 2) * (n_mels + 2) * sizeof(float);
        rc = launch_pdl(cnn14_conv1_kernel<float>, grid, dim3(kC1Threads), smem, st, lms, (const float*)net->bn0_s,
                        (const float*)net->bn0_b, (const float*)net->w1, (const float*)net->s1, (const float*)net->b1, cur,
                        n_mels, n_frames);
        if (rc) return rc;
        AC_LAUNCHED("cnn14_conv1_kernel");
    }
    int l = 0;
    for (int i = 0; i < kCnn14Blocks; ++i) {
        for (int j = 0; j < 2; ++j) {
            if (i == 0 && j == 0) continue;
            const Cnn14Conv& c = net->conv[l++];
            Conv3Args a; a.in = cur; a.out = nxt; a.B = B; a.H = d[i].H; a.W = d[i].W; a.Cin = c.cin; a.Cout = c.cout;
            a.bias = c.bias; a.tw = &c.tw; a.act = ACT_RELU; a.passes = net->conv_passes;
            rc = conv3x3_tc(a, st); if (rc) return rc;
            std::swap(cur, nxt);
        }
        if (i + 1 < kCnn14Blocks) {
            const int C4 = kCnn14Ch[i + 1] / 4, Ho = d[i].H / 2, Wo = d[i].W / 2;
            const int64_t total = (int64_t)B * Ho * Wo * C4;
            AC_TIMED("cnn14_avgpool", st);
            rc = launch_pdl(cnn14_avgpool_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st,
                            (const float4*)cur, (float4*)nxt, d[i].H, d[i].W, C4, Ho, Wo, total, dpc, kSiteCnn + (uint32_t)i);
            if (rc) return rc;
            AC_LAUNCHED("cnn14_avgpool_kernel");
            std::swap(cur, nxt);
        }
    }
    const Cnn14Dims last = d[kCnn14Blocks - 1];
    const int D = ac_cnn14_out_dim();
    rc = dropout_apply(cur, 0, (int64_t)B * last.H * last.W * D, dpc, kSiteCnn + kCnn14Blocks - 1, st); if (rc) return rc;
    {
        AC_TIMED("cnn14_tail", st);
        rc = launch_pdl(cnn14_tail_kernel<float>, dim3(cdiv(D, 128), B), dim3(128), 0, st, (const float*)cur, lens, attn_emb, pooled,
                        last.H, last.W, D);
        if (rc) return rc;
        AC_LAUNCHED("cnn14_tail_kernel");
    }
    rc = dropout_apply(pooled, 0, (int64_t)B * D, dpf, kSiteCnn + kCnn14Blocks, st); if (rc) return rc;
    GemmArgs g; g.A = pooled; g.W = net->fc_w; g.C = fc_emb; g.M = B; g.N = D; g.K = D; g.cbias = net->fc_b; g.act = ACT_RELU;
    g.tw = &net->fc_tw;
    rc = gemm_tn(g, st); if (rc) return rc;
    return dropout_apply(fc_emb, 0, (int64_t)B * D, dpf, kSiteCnn + kCnn14Blocks + 1, st);
}

}  // extern "C"

// =====================================================================================================================
// CNN8 + bi-GRU sound-event tagger of the temporal captioner (SED): captioning/models/hf_wrapper.py:1791-1859
// `Cnn8rnnSedModel.forward_prob` + the double threshold of :123-162.  Same building blocks as Cnn14: bn0 + first
// convolution (SIMT), 7 tensor-core 3x3 convolutions, 'avg+max' pooling (2,2) (2,2) (1,2) (1,2), mean over mel,
// fc1 + ReLU, one bidirectional GRU layer over all T/4 segments (csrc/bigru.cu), fc_audioset + sigmoid + clamp.
// The reference then repeat-upsamples x4 to frames, pads to the input length and thresholds on the CPU in numpy; here
// the hysteresis runs on the device at SEGMENT resolution (identical decisions: the upsampling is a pure repeat), and
// only the 0/1 label matrix goes back to the host for the tiny pairwise segment rule.
struct ac_sed {
    float* blob = nullptr;
    void* blob_bf16 = nullptr;    // bf16 images of the convolution weights (precision mode 16)
    float *bn0_s, *bn0_b, *w1, *s1, *b1;
    ac::Cnn14Conv conv[7];
    float *fc1_w, *fc1_b, *fco_w, *fco_b;
    ac::TcWeight fc1_tw, fco_tw;
    ac_bigru_t* gru = nullptr;
    int classes = 0, classes_pad = 0;
    int conv_passes = 3;
};

namespace ac {
constexpr int kSedBlocks = 4;
constexpr int kSedCh[kSedBlocks + 1] = {1, 64, 128, 256, 512};
constexpr int kSedPoolH[kSedBlocks] = {2, 2, 1, 1};
static void sed_walk(int n_mels, int n_frames, Cnn14Dims (&d)[kSedBlocks + 1]) {   // input dims of each block, then the output
    int H = n_frames, W = n_mels;
    for (int i = 0; i < kSedBlocks; ++i) { d[i] = {H, W}; H /= kSedPoolH[i]; W /= 2; }
    d[kSedBlocks] = {H, W};
}
static size_t sed_act_elems(int batch, int n_mels, int n_frames) {
    Cnn14Dims d[kSedBlocks + 1];
    sed_walk(n_mels, n_frames, d);
    size_t m = 0;
    for (int i = 0; i < kSedBlocks; ++i) m = std::max(m, (size_t)d[i].H * d[i].W * kSedCh[i + 1]);
    return align_up(m * batch, 64);
}
}  // namespace ac

extern "C" {

int ac_sed_set_precision(ac_sed_t* net, int tf32_passes) {
    using namespace ac;
    AC_REQUIRE(net && (tf32_passes == 1 || tf32_passes == 3 || tf32_passes == 16),
               "ac_sed_set_precision: mode must be 3 (3xTF32), 1 (TF32) or 16 (bf16)");
    net->conv_passes = tf32_passes;
    return AC_OK;
}

int ac_sed_num_tensors(void) { return 4 + ac::kSedBlocks * 10 + 2 + 8 + 2; }
int ac_sed_segments(int n_frames) { return n_frames / 4; }

int ac_sed_create(const float* const* t, const int64_t* numels, int n_tensors, int classes, void* stream, ac_sed_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_sed_create: null argument");
    AC_REQUIRE(n_tensors == ac_sed_num_tensors(), "ac_sed_create: expected %d tensors, got %d", ac_sed_num_tensors(), n_tensors);
    AC_REQUIRE(classes >= 1, "ac_sed_create: bad class count %d", classes);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = kSedCh[kSedBlocks], CP = (classes + 15) / 16 * 16;
    std::vector<int64_t> want;
    for (int q = 0; q < 4; ++q) want.push_back(64);
    for (int i = 0; i < kSedBlocks; ++i) {
        const int ci = kSedCh[i], co = kSedCh[i + 1];
        want.push_back((int64_t)co * ci * 9); want.push_back((int64_t)co * co * 9);
        for (int q = 0; q < 8; ++q) want.push_back(co);
    }
    want.push_back((int64_t)D * D); want.push_back(D);
    for (int d = 0; d < 2; ++d) { want.push_back(768 * (int64_t)D); want.push_back(768 * 256); want.push_back(768); want.push_back(768); }
    want.push_back((int64_t)classes * D); want.push_back(classes);
    for (int i = 0; i < n_tensors; ++i)
        AC_REQUIRE(numels[i] == want[i], "ac_sed_create: tensor %d has %lld elements, expected %lld", i, (long long)numels[i],
                   (long long)want[i]);
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += align_up(n, 64); return o; };
    const size_t o_bn0s = take(64), o_bn0b = take(64), o_w1 = take(9 * 64), o_s1 = take(64), o_b1 = take(64);
    struct Off { size_t s, b, pk; };
    Off co_[7]; int cin_[7], cout_[7];
    size_t perm_max = 0;
    {
        int l = 0;
        for (int i = 0; i < kSedBlocks; ++i)
            for (int j = 0; j < 2; ++j) {
                if (i == 0 && j == 0) continue;
                cin_[l] = j == 0 ? kSedCh[i] : kSedCh[i + 1]; cout_[l] = kSedCh[i + 1];
                co_[l].s = take(cout_[l]); co_[l].b = take(cout_[l]); co_[l].pk = take(tc_packed_floats(cout_[l], 9 * cin_[l]));
                perm_max = std::max(perm_max, (size_t)cout_[l] * 9 * cin_[l]);
                ++l;
            }
    }
    const size_t o_fc1w = take((size_t)D * D), o_fc1b = take(D), o_fc1pk = take(tc_packed_floats(D, D));
    const size_t o_fcow = take((size_t)CP * D), o_fcob = take(CP), o_fcopk = take(tc_packed_floats(CP, D));
    ac_sed_t* net = new ac_sed_t();
    net->classes = classes; net->classes_pad = CP;
    float* perm = nullptr;
    size_t bf_total = 0, bf_off[7];
    for (int q = 0; q < 7; ++q) { bf_off[q] = bf_total; bf_total += align_up(conv_bf16_packed_elems(cout_[q], cin_[q]), 64); }
    int rc = check_cuda(cudaMalloc(&net->blob, total * sizeof(float)), "ac_sed_create: cudaMalloc(weights)");
    if (rc == AC_OK) rc = check_cuda(cudaMalloc(&net->blob_bf16, bf_total * 2), "ac_sed_create: cudaMalloc(bf16 weights)");
    if (rc == AC_OK) rc = check_cuda(cudaMalloc(&perm, perm_max * sizeof(float)), "ac_sed_create: cudaMalloc(scratch)");
    if (rc == AC_OK) rc = check_cuda(cudaMemsetAsync(net->blob, 0, total * sizeof(float), st), "memset");
    if (rc != AC_OK) { cudaFree(net->blob); delete net; return rc; }
    float* B0 = net->blob;
    auto fold = [&](int ti, int c, float* s, float* b) {
        cnn14_bn_fold_kernel<<<cdiv(c, 256), 256, 0, st>>>(t[ti], t[ti + 1], t[ti + 2], t[ti + 3], kCnn14BnEps, s, b, c);
        g_launches++;
    };
    net->bn0_s = B0 + o_bn0s; net->bn0_b = B0 + o_bn0b; net->w1 = B0 + o_w1; net->s1 = B0 + o_s1; net->b1 = B0 + o_b1;
    fold(0, 64, net->bn0_s, net->bn0_b);
    int l = 0;
    for (int i = 0; i < kSedBlocks && rc == AC_OK; ++i) {
        const int base = 4 + i * 10;
        for (int j = 0; j < 2 && rc == AC_OK; ++j) {
            const int bn_ti = base + 2 + 4 * j;
            if (i == 0 && j == 0) {
                cnn14_w1_kernel<<<cdiv(9 * 64, 256), 256, 0, st>>>(t[base], net->w1, 64);
                g_launches++;
                fold(bn_ti, 64, net->s1, net->b1);
                continue;
            }
            Cnn14Conv& c = net->conv[l];
            c.cin = cin_[l]; c.cout = cout_[l]; c.scale = B0 + co_[l].s; c.bias = B0 + co_[l].b;
            fold(bn_ti, c.cout, c.scale, c.bias);
            rc = conv3x3_permute_weight(t[base + j], perm, c.cout, c.cin, st);
            if (rc == AC_OK) rc = tc_pack_weight(perm, c.scale, c.cout, 9 * c.cin, B0 + co_[l].pk, st, &c.tw);
            if (rc == AC_OK) rc = conv_bf16_pack(perm, c.scale, c.cout, c.cin, (uint16_t*)net->blob_bf16 + bf_off[l], st, &c.bw);
            ++l;
        }
    }
    int ti = 4 + kSedBlocks * 10;
    auto copy = [&](float* dst, const float* src, size_t n) {
        if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st), "ac_sed_create copy");
    };
    net->fc1_w = B0 + o_fc1w; net->fc1_b = B0 + o_fc1b; net->fco_w = B0 + o_fcow; net->fco_b = B0 + o_fcob;
    copy(net->fc1_w, t[ti], (size_t)D * D); copy(net->fc1_b, t[ti + 1], D);
    if (rc == AC_OK) rc = tc_pack_weight(net->fc1_w, nullptr, D, D, B0 + o_fc1pk, st, &net->fc1_tw);
    if (rc == AC_OK) rc = ac_bigru_create(t + ti + 2, numels + ti + 2, 8, D, 256, 1, stream, &net->gru);
    copy(net->fco_w, t[ti + 10], (size_t)classes * D); copy(net->fco_b, t[ti + 11], classes);   // rows >= classes stay zero
    if (rc == AC_OK) rc = tc_pack_weight(net->fco_w, nullptr, CP, D, B0 + o_fcopk, st, &net->fco_tw);
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "ac_sed_create pack kernels");
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_sed_create sync");
    cudaFree(perm);
    if (rc != AC_OK) { ac_bigru_destroy(net->gru); cudaFree(net->blob); delete net; return rc; }
    *out = net;
    return AC_OK;
}

void ac_sed_destroy(ac_sed_t* net) {
    if (!net) return;
    ac_bigru_destroy(net->gru);
    cudaFree(net->blob);
    cudaFree(net->blob_bf16);
    delete net;
}

size_t ac_sed_workspace_bytes(const ac_sed_t* net, int batch, int n_mels, int n_frames) {
    if (!net) return 0;
    using namespace ac;
    const int S = ac_sed_segments(n_frames), D = kSedCh[kSedBlocks];
    const size_t rows = (size_t)batch * S;
    return (2 * sed_act_elems(batch, n_mels, n_frames) + 3 * align_up(rows * D, 64) + align_up(rows * net->classes_pad, 64)) * sizeof(float) +
           align_up((size_t)batch * sizeof(int64_t), 256) + ac_bigru_workspace_bytes(net->gru, batch, S);
}

int ac_sed_fwd(const ac_sed_t* net, const float* lms, int B, int n_mels, int n_frames, float high, float low, float* prob_out,
               unsigned char* labels_out, int* runs_out, int max_runs, int* n_runs_out, void* workspace, size_t ws_bytes,
               void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 0 && B <= 65535, "ac_sed_fwd: batch %d out of range", B);
    AC_REQUIRE(n_mels == 64 && n_frames >= 16, "ac_sed_fwd: expects 64 mel bins and >= 16 frames (got %d x %d)", n_mels, n_frames);
    if (B == 0) return AC_OK;
    AC_REQUIRE(net && lms && labels_out, "ac_sed_fwd: null argument");
    AC_REQUIRE(workspace && ws_bytes >= ac_sed_workspace_bytes(net, B, n_mels, n_frames), "ac_sed_fwd: workspace too small (%zu < %zu)",
               ws_bytes, ac_sed_workspace_bytes(net, B, n_mels, n_frames));
    cudaStream_t st = (cudaStream_t)stream;
    const int S = ac_sed_segments(n_frames), D = kSedCh[kSedBlocks], CP = net->classes_pad;
    const size_t rows = (size_t)B * S;
    const size_t act = sed_act_elems(B, n_mels, n_frames);
    float* cur = (float*)workspace;
    float* nxt = cur + act;
    float* X = nxt + act;                        // [B, S, D] mean over mel
    float* F1 = X + align_up(rows * D, 64);      // relu(fc1)
    float* G = F1 + align_up(rows * D, 64);      // GRU output [B, S, 512]
    float* P = G + align_up(rows * D, 64);       // logits / probabilities [B, S, CP]
    int64_t* lens = (int64_t*)(P + align_up(rows * CP, 64));
    char* gws = (char*)lens + align_up((size_t)B * sizeof(int64_t), 256);
    Cnn14Dims d[kSedBlocks + 1];
    sed_walk(n_mels, n_frames, d);
    AC_REQUIRE(d[kSedBlocks].H == S, "ac_sed_fwd: internal frame count mismatch");
    int rc;
    if (net->conv_passes == 16) {
        // bf16 precision mode: activations in bf16 NHWC up to the mean over mel (see ac_cnn14_fwd_train)
        __nv_bfloat16* bcur = reinterpret_cast<__nv_bfloat16*>(cur);
        __nv_bfloat16* bnxt = reinterpret_cast<__nv_bfloat16*>(nxt);
        {
            AC_TIMED("sed_conv1", st);
            dim3 grid(cdiv(n_frames, kC1Time), B);
            const size_t smem = (size_t)(kC1Time + 2) * (n_mels + 2) * sizeof(float);
            rc = launch_pdl(cnn14_conv1_kernel<__nv_bfloat16>, grid, dim3(kC1Threads), smem, st, lms, (const float*)net->bn0_s,
                            (const float*)net->bn0_b, (const float*)net->w1, (const float*)net->s1, (const float*)net->b1, bcur,
                            n_mels, n_frames);
            if (rc) return rc;
            AC_LAUNCHED("cnn14_conv1_kernel");
        }
        int l = 0;
        for (int i = 0; i < kSedBlocks; ++i) {
            for (int j = 0; j < 2; ++j) {
                if (i == 0 && j == 0) continue;
                const Cnn14Conv& c = net->conv[l++];
                ConvBf16Args a; a.in = bcur; a.out = bnxt; a.B = B; a.H = d[i].H; a.W = d[i].W; a.Cin = c.cin; a.Cout = c.cout;
                a.bias = c.bias; a.w = &c.bw; a.act = ACT_RELU;
                rc = conv3x3_bf16(a, st); if (rc) return rc;
                std::swap(bcur, bnxt);
            }
            const int C8 = kSedCh[i + 1] / 8, Ho = d[i + 1].H, Wo = d[i + 1].W;
            const int64_t total = (int64_t)B * Ho * Wo * C8;
            AC_TIMED("sed_pool", st);
            rc = launch_pdl(cnn_avgmax_pool_bf16_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, (const uint4*)bcur,
                            (uint4*)bnxt, d[i].H, d[i].W, C8, Ho, Wo, kSedPoolH[i], 2, total);
            if (rc) return rc;
            AC_LAUNCHED("cnn_avgmax_pool_kernel");
            std::swap(bcur, bnxt);
        }
        const int64_t total = (int64_t)rows * (D / 8);
        rc = launch_pdl(cnn_wmean_bf16_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, (const uint4*)bcur, (float4*)X,
                        d[kSedBlocks].W, D / 8, total);
        if (rc) return rc;
        AC_LAUNCHED("cnn_wmean_kernel");
    } else {
    {
        AC_TIMED("sed_conv1", st);
        dim3 grid(cdiv(n_frames, kC1Time), B);
        const size_t smem = (size_t)(kC1Time + 2) * (n_mels + 2) * sizeof(float);
        rc = launch_pdl(cnn14_conv1_kernel<float>, grid, dim3(kC1Threads), smem, st, lms, (const float*)net->bn0_s, (const float*)net->bn0_b,
                        (const float*)net->w1, (const float*)net->s1, (const float*)net->b1, cur, n_mels, n_frames);
        if (rc) return rc;
        AC_LAUNCHED("cnn14_conv1_kernel");
    }
    int l = 0;
    for (int i = 0; i < kSedBlocks; ++i) {
        for (int j = 0; j < 2; ++j) {
            if (i == 0 && j == 0) continue;
            const Cnn14Conv& c = net->conv[l++];
            Conv3Args a; a.in = cur; a.out = nxt; a.B = B; a.H = d[i].H; a.W = d[i].W; a.Cin = c.cin; a.Cout = c.cout;
            a.bias = c.bias; a.tw = &c.tw; a.act = ACT_RELU; a.passes = net->conv_passes;
            rc = conv3x3_tc(a, st); if (rc) return rc;
            std::swap(cur, nxt);
        }
        const int C4 = kSedCh[i + 1] / 4, Ho = d[i + 1].H, Wo = d[i + 1].W;
        const int64_t total = (int64_t)B * Ho * Wo * C4;
        AC_TIMED("sed_pool", st);
        rc = launch_pdl(cnn_avgmax_pool_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, (const float4*)cur, (float4*)nxt,
                        d[i].H, d[i].W, C4, Ho, Wo, kSedPoolH[i], 2, total);
        if (rc) return rc;
        AC_LAUNCHED("cnn_avgmax_pool_kernel");
        std::swap(cur, nxt);
    }
    {
        const int64_t total = (int64_t)rows * (D / 4);
        rc = launch_pdl(cnn_wmean_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, (const float4*)cur, (float4*)X,
                        d[kSedBlocks].W, D / 4, total);
        if (rc) return rc;
        AC_LAUNCHED("cnn_wmean_kernel");
    }
    }
    GemmArgs g; g.A = X; g.W = net->fc1_w; g.C = F1; g.M = (int)rows; g.N = D; g.K = D; g.cbias = net->fc1_b; g.act = ACT_RELU; g.tw = &net->fc1_tw;
    rc = gemm_tn(g, st); if (rc) return rc;
    // the SED GRU runs over the full sequence of every clip (no packing): lens = S
    sed_fill_kernel<<<cdiv(B, 256), 256, 0, st>>>(lens, (int64_t)S, B);
    AC_LAUNCHED("sed_fill_kernel");
    rc = ac_bigru_fwd(net->gru, F1, lens, B, S, S, G, gws, ac_bigru_workspace_bytes(net->gru, B, S), stream); if (rc) return rc;
    GemmArgs go; go.A = G; go.W = net->fco_w; go.C = P; go.M = (int)rows; go.N = CP; go.K = D; go.cbias = net->fco_b; go.act = ACT_NONE; go.tw = &net->fco_tw;
    rc = gemm_tn(go, st); if (rc) return rc;
    {
        const int64_t total = (int64_t)rows * CP;
        rc = launch_pdl(sed_sigmoid_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, st, P, total);
        if (rc) return rc;
        AC_LAUNCHED("sed_sigmoid_kernel");
        const int tot = B * net->classes;
        if (n_runs_out != nullptr) AC_CUDA(cudaMemsetAsync(n_runs_out, 0, sizeof(int), st));
        AC_TIMED("sed_hysteresis", st);
        rc = launch_pdl(sed_hysteresis_kernel, dim3(cdiv(tot, 128)), dim3(128), 0, st, (const float*)P, labels_out, S, CP, net->classes,
                        high, low, tot, (int4*)runs_out, runs_out ? max_runs : 0, runs_out ? n_runs_out : (int*)nullptr);
        if (rc) return rc;
        AC_LAUNCHED("sed_hysteresis_kernel");
    }
    if (prob_out != nullptr)   // [B, S, classes] (un-padded copy for inspection / tests)
        AC_CUDA(cudaMemcpy2DAsync(prob_out, (size_t)net->classes * sizeof(float), P, (size_t)CP * sizeof(float),
                                  (size_t)net->classes * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
    return AC_OK;
}

}  // extern "C"
