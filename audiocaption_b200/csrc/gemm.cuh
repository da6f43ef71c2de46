// fp32 "pointwise" GEMM used by every 1x1 convolution / linear layer on the path:
//     C[m, n] = act( (sum_k A[m,k] * ascale[m / rows_per_group, k] * W[n,k]) * cscale[n] + cbias[n] ) + R[m, n]
// A is [M, K] row-major (NHWC activations: K = input channels), W is [N, K] row-major
// (exactly the layout of a PyTorch Conv2d 1x1 / Linear weight), so both operands stream
// along K with 128-bit loads.  Tiles are staged K-major in shared memory with a
// register-prefetch double buffer.  K % 4 == 0 and N % 4 == 0 are required (true for every
// layer of EfficientNet-B2 and of the caption decoder).
#pragma once
#include "common.cuh"

namespace ac {

enum Act { ACT_NONE = 0, ACT_SWISH = 1, ACT_RELU = 2 };

// A weight matrix W[N, K] re-packed for the tcgen05 kernel (gemm_tc.cu): per (n-tile, 32-wide k-chunk)
// one contiguous block holding the tf32 "hi" part and the fp32 residual "lo" part, each already in the
// 128B-swizzled K-major shared-memory image the tensor core reads, so a CTA fetches a chunk with a
// single bulk copy.  Rows/columns beyond N/K are zero.
struct TcWeight {
    const float* packed = nullptr;
    const float* scale = nullptr;   // per-output-channel scale folded into the packed copy (nullptr = none)
    int N = 0, K = 0;
    int BN = 0;        // columns per n-tile (multiple of 16, <= 160)
    int n_tiles = 0;
    int k_chunks = 0;  // ceil(K / 32)
};
size_t tc_packed_floats(int N, int K);
// Narrow images (n-tiles of at most bn_cap columns, streamed): the same GEMM cut into more CTAs, for launches whose M is
// too small to fill the SMs with the default tiling.
size_t tc_packed_floats_bn(int N, int K, int bn_cap);
int tc_pack_weight_bn(const float* W_dev, const float* scale_dev, int N, int K, int bn_cap, float* dst_dev, cudaStream_t st,
                      TcWeight* out);
// W_dev [N, K] row-major -> dst_dev (tc_packed_floats(N, K) floats); fills *out.
// scale_dev [N] (nullable) is multiplied into the rows before the hi/lo split (folded BN scale).
int tc_pack_weight(const float* W_dev, const float* scale_dev, int N, int K, float* dst_dev, cudaStream_t st,
                   TcWeight* out);

// Same, for a matrix seen through strides: element (n, k) is W_dev[n * sn + k * sk] (sn = 1, sk = ld packs the
// TRANSPOSE of a row-major matrix: the backward GEMMs of the training path, train_ops.cu).
int tc_pack_weight_strided(const float* W_dev, const float* scale_dev, int N, int K, int64_t sn, int64_t sk, float* dst_dev,
                           cudaStream_t st, TcWeight* out);

// Batched packing: a static table of jobs (one per matrix image) lives in device memory; tc_pack_multi re-packs them all
// in ONE launch.  tc_pack_plan fills a job + its TcWeight and returns the job's item count (float4 units); `first` is the
// running sum of the counts of the jobs before it.
struct TcPackJob { const float* W; float* dst; int N, K, BN, n_tiles, k_chunks; long long sn, sk, first; };
long long tc_pack_plan(const float* W, int N, int K, long long sn, long long sk, float* dst, long long first, TcPackJob* job,
                       TcWeight* out);
int tc_pack_multi(const TcPackJob* jobs_dev, int n_jobs, long long total_items, cudaStream_t st);

struct GemmArgs {
    const float* A; const float* W; float* C;
    int M, N, K;
    const float* ascale = nullptr;  // [M / rows_per_group, K] per-group input-channel scale (SE gate)
    int rows_per_group = 1;
    const float* cscale = nullptr;  // [N] (folded BN gamma/sqrt(var+eps)); nullptr = 1
    const float* cbias = nullptr;   // [N]; nullptr = 0
    const float* R = nullptr;       // [M, N] residual added after the activation
    int act = ACT_NONE;
    int ldc = 0;                    // row stride of C / R (0 -> N)
    int lda = 0;                    // row stride of A (0 -> K); a multiple of 4 floats
    int passes = 3;                 // tensor-core path: 3 = 3xTF32 (fp32-level accuracy), 1 = plain TF32
    const TcWeight* tw = nullptr;   // packed copy of W for the tensor-core path (nullptr -> SIMT kernel)
};

// Dispatch: the tcgen05 kernel (3xTF32 split, fp32-level accuracy) when g.tw is set and K % 8 == 0,
// otherwise the plain-fp32 SIMT kernel.
int gemm_tn(const GemmArgs& g, cudaStream_t st);
int gemm_tn_simt(const GemmArgs& g, cudaStream_t st);
int gemm_tc(const GemmArgs& g, cudaStream_t st);
void* tensor_map_encode_fn();   // PFN_cuTensorMapEncodeTiled or nullptr

// 3x3 / stride 1 / pad 1 convolution + bias + activation on the tensor-core pipeline (gemm_tc.cu), NHWC fp32.
// tw packs the [Cout, 9*Cin] tap-major weight made by conv3x3_permute_weight (BN scale folded by tc_pack_weight).
struct Conv3Args {
    const float* in; float* out;
    int B, H, W, Cin, Cout;
    const float* bias = nullptr;    // [Cout]
    const TcWeight* tw = nullptr;
    int act = ACT_NONE;
    int passes = 3;                 // 3 = 3xTF32, 1 = plain TF32 ("tf32" precision mode)
};
int conv3x3_tc(const Conv3Args& a, cudaStream_t st);
int conv3x3_permute_weight(const float* w_dev, float* out_dev, int Cout, int Cin, cudaStream_t st);
void conv3x3_tile_shape(int H, int W, int& Hbox, int& Bbox);

// The same convolution with bf16 activations and weights (conv_bf16.cu): in / out NHWC bf16, bias fp32; the weight is the
// tap-major [Cout, 9*Cin] matrix (conv3x3_permute_weight) packed to bf16 with the BN scale folded in.
struct ConvBf16Weight { const void* packed = nullptr; int Cout = 0, Cin = 0, BN = 0, n_tiles = 0, k_chunks = 0; };
struct ConvBf16Args {
    const void* in; void* out; const float* bias; const ConvBf16Weight* w = nullptr;
    int B, H, W, Cin, Cout;
    int act = ACT_NONE;
    int max_ctas = 0;            // persistent CTAs of the launch (0 = one per SM): ac_cnn14_set_sm_limit
};
size_t conv_bf16_packed_elems(int Cout, int Cin);
int conv_bf16_pack(const float* w_perm_dev, const float* scale_dev, int Cout, int Cin, void* dst_dev, cudaStream_t st,
                   ConvBf16Weight* out);
int conv3x3_bf16(const ConvBf16Args& a, cudaStream_t st);

// Depthwise k x k convolution + folded BN + swish + per-tile channel sums (SE squeeze), NHWC fp32, fed by 4-D
// TMA tiles (dwconv_tma.cu).  in [B,Hi,Wi,C] -> out [B,Ho,Wo,C]; partial [B][dwconv_tiles_per_clip][C].
struct DwArgs {
    const float* in; float* out; float* partial;
    const float* w;       // [k*k][C]
    const float* scale;   // [C] folded BN
    const float* bias;    // [C]
    int B, Hi, Wi, Ho, Wo, C, k, s, pad_lo;
};
int dwconv_tiles_per_clip(int Ho, int Wo, int C, int k, int s);   // second dimension of `partial` (tiles x warps)
int dwconv_tma(const DwArgs& a, cudaStream_t st);

}  // namespace ac
