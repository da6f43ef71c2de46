// 3x3 / stride 1 / pad 1 convolution + folded BN + ReLU with bf16 activations and weights (sm_100a): the "bf16" precision
// mode of the frozen Cnn14 encoder (captioning/models/cnn_encoder.py:32-75 `ConvBlock`, eval-mode BatchNorm), for the
// configurations BASELINE.json quotes in bf16 (training and the GRU-attention captioner).
//
//   out[b,h,w,n] = relu( sum_{ky,kx,c} in[b, h+ky-1, w+kx-1, c] * Wp[n, (ky*3+kx)*Cin + c] + bias[n] )      NHWC, bf16 in / out
//
// Why a separate kernel: the fp32 / TF32 implicit GEMM (gemm_tc.cu) is bound by what one SM can pull from L2 per k-chunk
// (16 KB of fp32 activations + 16..32 KB of tf32 weights per 32 input channels; DESIGN.md "Measured and rejected"), not by
// the tensor pipe.  With bf16 operands a 128-byte swizzled row holds 64 channels, so the same 32 KB stage covers twice the
// K, both operands go from shared memory straight into `tcgen05.mma kind::f16` (no register/TMEM transform stage), and
// the layer outputs are written as bf16, which halves the activation traffic of the next layer as well.
//
// Structure (one persistent CTA per SM, 6 warps):
//   warp 0     TMA producer: per 64-channel k-chunk (tap, channel block) one 4-D box of the NHWC input shifted by the tap
//              (out-of-image rows / columns arrive as zeros = the padding) + one bulk copy of the packed weight chunk
//   warp 1     MMA issuer: 4 x tcgen05.mma (M = 128 pixels, N = BN, K = 16) per chunk into one of two TMEM accumulators;
//              tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld, + bias, ReLU, pack to bf16, 16-byte stores of each pixel's channel run
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace ac {

constexpr int CB_BM = 128;
constexpr int CB_KC = 64;                          // bf16 channels per k-chunk = one 128-byte swizzle row
constexpr int CB_A_BYTES = CB_BM * 128;            // 16 KB
// Widest n-tile.  A k-chunk of 64 channels brings 16 KB of activations + BN x 128 B of weights into the SM for
// 128 x BN x 64 MACs: at BN = 128 that is 32 KB per ~262 tensor cycles = 125 B/clk, twice what one SM pulls from L2
// (~60 B/clk; ncu: tensor pipe 54 % busy) -- at BN = 256 the same activations serve twice the columns: 48 KB per ~524
// cycles = 92 B/clk.  AC_CONV_BF16_BN=128 restores the narrow tiles (read when the weights are packed).
constexpr int CB_MAX_BN_LIMIT = 256;
static int cb_max_bn() {
    static const int v = [] { const char* e = getenv("AC_CONV_BF16_BN"); const int x = e ? atoi(e) : 256;
                              return x == 128 || x == 256 ? x : 256; }();
    return v;
}
constexpr int CB_THREADS = 192;
constexpr int CB_MAX_STAGES = 6;
constexpr int CB_SMEM_LIMIT = 227 * 1024;

struct ConvBf16Params {
    const __nv_bfloat16* wpacked; __nv_bfloat16* out; const float* bias;
    int N, BN, n_tiles, m_tiles, k_chunks, stages, tmem_cols, act;
    int cW, cH, cB, cHbox, cBbox, c_tiles_h, c_cpc, a_bytes;
};

__global__ void __launch_bounds__(CB_THREADS, 1)
conv3x3_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const ConvBf16Params p) {
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;          // 128B swizzle needs 1024-byte aligned tiles
    const uint32_t w_bytes = (uint32_t)p.BN * 128u;
    const uint32_t stage_bytes = CB_A_BYTES + w_bytes;
    const uint32_t bars = base + p.stages * stage_bytes;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 64u + 8u * s; };
    auto bar_acc_full = [&](int a) { return bars + 128u + 8u * a; };
    auto bar_acc_empty = [&](int a) { return bars + 144u + 8u * a; };
    const uint32_t tmem_slot_addr = bars + 160u;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot_addr - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapA);
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), 4); }
        fence_mbar_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot_addr, (uint32_t)p.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();                                      // the input is the previous kernel's output

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_t = tile % p.n_tiles, m_t = tile / p.n_tiles;
                const __nv_bfloat16* wsrc = p.wpacked + (size_t)n_t * p.k_chunks * (p.BN * CB_KC);
                const int b0 = (m_t / p.c_tiles_h) * p.cBbox, h0 = (m_t % p.c_tiles_h) * p.cHbox;
                int tap = 0, cc = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(bar_empty(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * stage_bytes;
                    mbar_expect_tx(bar_full(stage), (uint32_t)p.a_bytes + w_bytes);
                    tma_load_4d(sa, &mapA, cc * CB_KC, tap % 3 - 1, h0 + tap / 3 - 1, b0, bar_full(stage));
                    bulk_load(sa + CB_A_BYTES, wsrc + (size_t)kc * (p.BN * CB_KC), w_bytes, bar_full(stage));
                    if (++cc == p.c_cpc) { cc = 0; ++tap; }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc(1 /* bf16 */, CB_BM, p.BN);
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(bar_acc_empty(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                mbar_wait(bar_full(stage), phase);
                tc_fence_after();
                if (leader) {
                    const uint32_t sa = base + stage * stage_bytes;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + CB_A_BYTES);
#pragma unroll
                    for (int ks = 0; ks < CB_KC / 16; ++ks)      // 16 bf16 = 32 bytes along the swizzled row = +2 in the address field
                        mma_bf16(d_tmem, da + 2u * ks, db + 2u * ks, idesc, (kc | ks) != 0 ? 1u : 0u);
                    mma_commit(bar_empty(stage));
                    if (kc + 1 == p.k_chunks) mma_commit(bar_acc_full(acc));
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5: TMEM lane quarter = warp & 3)
        const int q = warp & 3;
        const int r = q * 32 + lane;                        // row of the tile = pixel
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_t = tile % p.n_tiles, m_t = tile / p.n_tiles;
            const int w_ = r % p.cW, t_ = r / p.cW;
            const int hl = t_ % p.cHbox, bl = t_ / p.cHbox;
            const int h_ = (m_t % p.c_tiles_h) * p.cHbox + hl, b_ = (m_t / p.c_tiles_h) * p.cBbox + bl;
            const bool ok = bl < p.cBbox && h_ < p.cH && b_ < p.cB;
            __nv_bfloat16* orow = p.out + (((size_t)b_ * p.cH + h_) * p.cW + w_) * p.N + n_t * p.BN;
            const float* bias = p.bias + n_t * p.BN;
            mbar_wait(bar_acc_full(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t u[2][16];
                tmem_ld16(t_row + (uint32_t)c0, u[0]);
                tmem_ld16(t_row + (uint32_t)(c0 + 16), u[1]);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * j));
                    float v0 = __uint_as_float(u[j >> 2][(j & 3) * 4 + 0]) + bq.x, v1 = __uint_as_float(u[j >> 2][(j & 3) * 4 + 1]) + bq.y;
                    float v2 = __uint_as_float(u[j >> 2][(j & 3) * 4 + 2]) + bq.z, v3 = __uint_as_float(u[j >> 2][(j & 3) * 4 + 3]) + bq.w;
                    if (p.act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
                    packed[2 * j] = *reinterpret_cast<const uint32_t*>(&lo);
                    packed[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&hi);
                }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(orow + c0 + 8 * j) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------ weight packing
// w_perm [Cout][9 * Cin] fp32, tap-major (conv3x3_permute_weight), scale [Cout] (folded BN) ->
// [n_tile][k_chunk of 64][BN rows][128 bytes, 16-byte pieces swizzled by (row & 7)] bf16.  One thread per 16-byte piece.
__global__ void conv_bf16_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst,
                                      int N, int K, int BN, int n_tiles, int k_chunks) {
    const int64_t total = (int64_t)n_tiles * k_chunks * BN * 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int pc = (int)(i % 8);
    int64_t r = i / 8;
    const int row = (int)(r % BN); r /= BN;
    const int kc = (int)(r % k_chunks);
    const int n_t = (int)(r / k_chunks);
    const int lc = pc ^ (row & 7);                      // logical piece stored at physical piece pc
    const int n = n_t * BN + row, k = kc * CB_KC + lc * 8;
    uint32_t o[4] = {0u, 0u, 0u, 0u};
    if (n < N) {
        const float sc = scale != nullptr ? scale[n] : 1.0f;
        for (int e = 0; e < 4; ++e) {
            const float a = k + 2 * e < K ? w[(size_t)n * K + k + 2 * e] * sc : 0.f;
            const float b = k + 2 * e + 1 < K ? w[(size_t)n * K + k + 2 * e + 1] * sc : 0.f;
            const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
            o[e] = *reinterpret_cast<const uint32_t*>(&v);
        }
    }
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

static void conv_bf16_tiling(int Cout, int Cin, int& BN, int& n_tiles, int& k_chunks) {
    BN = std::min(cb_max_bn(), Cout);
    n_tiles = cdiv(Cout, BN);
    k_chunks = 9 * Cin / CB_KC;
}

size_t conv_bf16_packed_elems(int Cout, int Cin) {
    int BN, n_tiles, k_chunks;
    conv_bf16_tiling(Cout, Cin, BN, n_tiles, k_chunks);
    return (size_t)n_tiles * k_chunks * BN * CB_KC;
}

int conv_bf16_pack(const float* w_perm_dev, const float* scale_dev, int Cout, int Cin, void* dst_dev, cudaStream_t st,
                   ConvBf16Weight* out) {
    AC_REQUIRE(w_perm_dev && dst_dev && out, "conv_bf16_pack: null argument");
    AC_REQUIRE(Cin % CB_KC == 0 && Cout % 32 == 0 && (Cout <= cb_max_bn() || Cout % cb_max_bn() == 0),
               "conv_bf16_pack: Cin (%d) %% 64, Cout (%d) %% 32 (and %% %d beyond %d) must be 0", Cin, Cout, cb_max_bn(), cb_max_bn());
    AC_REQUIRE(((uintptr_t)dst_dev & 127) == 0, "conv_bf16_pack: destination must be 128-byte aligned");
    int BN, n_tiles, k_chunks;
    conv_bf16_tiling(Cout, Cin, BN, n_tiles, k_chunks);
    const int64_t total = (int64_t)n_tiles * k_chunks * BN * 8;
    conv_bf16_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(w_perm_dev, scale_dev, (__nv_bfloat16*)dst_dev, Cout,
                                                                        9 * Cin, BN, n_tiles, k_chunks);
    AC_LAUNCHED("conv_bf16_pack_kernel");
    out->packed = dst_dev; out->Cout = Cout; out->Cin = Cin; out->BN = BN; out->n_tiles = n_tiles; out->k_chunks = k_chunks;
    return AC_OK;
}

int conv3x3_bf16(const ConvBf16Args& a, cudaStream_t st) {
    AC_REQUIRE(a.w != nullptr && a.w->packed != nullptr, "conv3x3_bf16: weight not packed");
    const ConvBf16Weight& w = *a.w;
    AC_REQUIRE(w.Cin == a.Cin && w.Cout == a.Cout, "conv3x3_bf16: packed weight is %dx%d, call wants %dx%d", w.Cout, w.Cin,
               a.Cout, a.Cin);
    AC_REQUIRE(a.W >= 1 && a.W <= CB_BM && a.H >= 1, "conv3x3_bf16: width %d must be in 1..128", a.W);
    AC_REQUIRE(((uintptr_t)a.in & 15) == 0 && ((uintptr_t)a.out & 15) == 0, "conv3x3_bf16: buffers must be 16-byte aligned");
    if (a.B <= 0) return AC_OK;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(tensor_map_encode_fn());
    AC_REQUIRE(encode != nullptr, "conv3x3_bf16: cuTensorMapEncodeTiled is not available from the driver");
    int Hbox, Bbox;
    conv3x3_tile_shape(a.H, a.W, Hbox, Bbox);
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.Cin * 2, (cuuint64_t)a.W * a.Cin * 2, (cuuint64_t)a.H * a.W * a.Cin * 2};
    const cuuint32_t box[4] = {CB_KC, (cuuint32_t)a.W, (cuuint32_t)Hbox, (cuuint32_t)Bbox};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AC_REQUIRE(cr == CUDA_SUCCESS, "conv3x3_bf16: cuTensorMapEncodeTiled failed (%d) for in=%p [%d,%d,%d,%d]", (int)cr, a.in,
               a.B, a.H, a.W, a.Cin);
    ConvBf16Params p;
    p.wpacked = (const __nv_bfloat16*)w.packed; p.out = (__nv_bfloat16*)a.out; p.bias = a.bias;
    p.N = a.Cout; p.BN = w.BN; p.n_tiles = w.n_tiles; p.k_chunks = w.k_chunks; p.act = a.act;
    p.cW = a.W; p.cH = a.H; p.cB = a.B; p.cHbox = Hbox; p.cBbox = Bbox; p.c_tiles_h = cdiv(a.H, Hbox); p.c_cpc = a.Cin / CB_KC;
    p.m_tiles = cdiv(a.B, Bbox) * p.c_tiles_h;
    p.a_bytes = a.W * Hbox * Bbox * 128;
    const int sb = CB_A_BYTES + w.BN * 128;
    const int fixed = 1024 + 256;
    p.stages = std::min(CB_MAX_STAGES, (CB_SMEM_LIMIT - fixed) / sb);
    p.tmem_cols = 2 * w.BN <= 32 ? 32 : (2 * w.BN <= 64 ? 64 : (2 * w.BN <= 128 ? 128 : (2 * w.BN <= 256 ? 256 : 512)));
    AC_REQUIRE(w.BN <= CB_MAX_BN_LIMIT && p.stages >= 2, "conv3x3_bf16: n-tile %d does not fit (stages %d)", w.BN, p.stages);
    const size_t smem = (size_t)p.stages * sb + fixed;
    const int grid = std::min(p.m_tiles * p.n_tiles, a.max_ctas > 0 ? std::min(a.max_ctas, kNumSMs) : kNumSMs);
    AC_TIMED("conv3x3_bf16", st);
    static cudaError_t attr_rc = cudaFuncSetAttribute(conv3x3_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM_LIMIT);
    AC_CUDA(attr_rc);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(CB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1] = {pdl_attr()};
    cfg.attrs = at; cfg.numAttrs = 1;
    AC_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_bf16_kernel, map, p));
    AC_LAUNCHED("conv3x3_bf16_kernel");
    return AC_OK;
}

}  // namespace ac

// Diagnostic / test entry: 3x3 pad-1 convolution + per-channel scale/bias + activation on bf16 NHWC tensors.
// in_dev / out_dev: bf16 [B,H,W,Cin] / [B,H,W,Cout]; w_dev: the PyTorch Conv2d weight [Cout, Cin, 3, 3] fp32; act 0 none, 2 relu.
extern "C" int ac_conv3x3_bf16(const void* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev,
                               void* out_dev, int B, int H, int W, int Cin, int Cout, int act, void* stream) {
    using namespace ac;
    AC_REQUIRE(in_dev && w_dev && out_dev && bias_dev, "ac_conv3x3_bf16: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    float* perm = nullptr; void* packed = nullptr;
    AC_CUDA(cudaMalloc(&perm, (size_t)Cout * 9 * Cin * sizeof(float)));
    int rc = conv3x3_permute_weight(w_dev, perm, Cout, Cin, st);
    ConvBf16Weight cw;
    if (rc == AC_OK) rc = check_cuda(cudaMalloc(&packed, conv_bf16_packed_elems(Cout, Cin) * 2), "ac_conv3x3_bf16: cudaMalloc");
    if (rc == AC_OK) rc = conv_bf16_pack(perm, scale_dev, Cout, Cin, packed, st, &cw);
    if (rc == AC_OK) {
        ConvBf16Args a; a.in = in_dev; a.out = out_dev; a.bias = bias_dev; a.w = &cw; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
        a.Cout = Cout; a.act = act;
        rc = conv3x3_bf16(a, st);
    }
    cudaStreamSynchronize(st);
    cudaFree(perm); cudaFree(packed);
    return rc;
}
