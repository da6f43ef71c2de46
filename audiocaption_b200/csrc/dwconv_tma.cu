// Depthwise convolution of the MBConv blocks, fed by TMA tiles (sm_100a).
//
//   out[b,ho,wo,c] = swish( bn( sum_{kh,kw} in[b, ho*s+kh-pad, wo*s+kw-pad, c] * w[kh,kw,c] ) )
//   partial[b, tile, c] = sum of out over the tile's pixels          (squeeze-and-excitation numerator)
//
// (efficientnet_pytorch 0.7.1 `_depthwise_conv` (Conv2dStaticSamePadding) + `_bn1` + swish + the
//  `adaptive_avg_pool2d` of the SE branch, as driven by captioning/models/hf_wrapper.py:218-241.)
//
// HBM-bound op: every input element is needed by up to k*k outputs but should be read from HBM once.  A
// persistent CTA walks (clip, row-block, column-block, 32-channel chunk) tiles; a producer warp keeps a ring of
// input tiles (halo included) in flight with cp.async.bulk.tensor (4-D tensor map over the NHWC activation;
// the "static same" zero padding is simply the TMA out-of-bounds fill, negative coordinates included), so
// ~100+ KB per SM are in flight without costing registers.  Compute threads own a channel PAIR of one output
// row and slide along W in scatter form: one input column (k rows, LDS.64) is added into the <= k outputs it
// touches, the k*k weights live in registers.  A pixel's 32-channel chunk is 128 contiguous bytes in shared
// memory, so a half-warp reads/writes whole 128-byte lines (conflict-free LDS, full-line global stores).
#include <cudaTypedefs.h>

#include <algorithm>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace ac {

// CC = channels per tile (16, 32 or 64: 64 / 128 / 256 contiguous bytes per pixel); a pixel's chunk is served by
// CC / CPT threads (CPT = channels per thread), so the compute threads form 256 / (CC/2) (row, column sub-segment) groups.
// CPT = 2 everywhere (a channel pair per thread, LDS.64 / STG.64, 256 compute threads; 3x3 kernels run two CTAs per SM).
// The 5x5 kernels hold 50 weight registers per thread and fit only one 9-warp CTA per SM; the CPT = 1 variant of the same
// code (512 compute threads, 96 registers, 17 warps per SM) was measured SLOWER (all 5x5 layers of a 64-clip step:
// 0.445 -> 0.493 ms): twice the LDS / STG instructions outweigh the extra latency hiding.  dw_cpt() keeps the switch.
constexpr int DW_GROUP_THREADS = 256;                        // (compute threads) x CPT / 2: fixes the tile shape
__host__ __device__ constexpr int dw_cpt(int k) { (void)k; return 2; }
__host__ __device__ constexpr int dw_compute_threads(int k) { return DW_GROUP_THREADS * 2 / dw_cpt(k); }
// partial-sum slots per tile: one per compute warp, or per group of warps that share a pixel chunk (CC / CPT > 32 threads)
__host__ __device__ constexpr int dw_slots(int k, int cc) { return (dw_compute_threads(k) / 32) / (cc / dw_cpt(k) > 32 ? cc / dw_cpt(k) / 32 : 1); }
constexpr int DW_MAX_STAGES = 4;
constexpr int DW_SMEM_LIMIT = 200 * 1024;

struct DwParams {
    float* out; float* partial; const float* w; const float* scale; const float* bias;
    int B, Ho, Wo, C, pad_lo;
    int CC, groups;
    int Ht, Ws, nsub, n_t;            // tile: Ht output rows x Ws output columns; nsub column sub-segments of n_t
    int tiles_h, tiles_w, chunks, total_tiles;
    int Hbox, Wbox, stages, tile_bytes;
};

using ptx::tma_load_4d;

// tile index -> (chunk, b, th, tw): chunk is the slowest so that a CTA's consecutive tiles share weights
__device__ __forceinline__ void dw_tile_coords(const DwParams& p, int tile, int& chunk, int& b, int& th, int& tw) {
    tw = tile % p.tiles_w; tile /= p.tiles_w;
    th = tile % p.tiles_h; tile /= p.tiles_h;
    b = tile % p.B;
    chunk = tile / p.B;
}

template <int K, int S, int CC>
__global__ void __launch_bounds__(dw_compute_threads(K) + 32, K == 3 ? 2 : 1)     // 3x3: 2 CTAs per SM (16 compute warps)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap mapIn, const DwParams p) {
    using namespace ptx;
    constexpr int CPT = dw_cpt(K);                       // channels per thread
    constexpr int NT = dw_compute_threads(K);            // compute threads (+ one producer warp)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    const uint32_t bars = base + p.stages * p.tile_bytes;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 32u + 8u * s; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        prefetch_tensormap(&mapIn);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), NT / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_trigger();
    pdl_wait();        // the input activation is the previous kernel's output

    if (warp == NT / 32) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int chunk, b, th, tw;
                dw_tile_coords(p, tile, chunk, b, th, tw);
                mbar_wait(bar_empty(stage), phase ^ 1u);
                mbar_expect_tx(bar_full(stage), (uint32_t)p.tile_bytes);
                tma_load_4d(base + stage * p.tile_bytes, &mapIn, chunk * CC, tw * p.Ws * S - p.pad_lo,
                            th * p.Ht * S - p.pad_lo, b, bar_full(stage));
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- compute threads
    constexpr int TPG = CC / CPT;            // threads per pixel chunk
    const int cp = tid % TPG;                // channel group inside the chunk
    const int g = tid / TPG;                 // group: output row r (fast) x column sub-segment
    const int r = g % p.Ht, sub = g / p.Ht;
    const int wl0 = sub * p.n_t;             // first local output column of this thread
    float wr[K * K][CPT];
    float sc[CPT], bi[CPT];
#pragma unroll
    for (int e = 0; e < CPT; ++e) { sc[e] = 0.f; bi[e] = 0.f; }
    int cur_chunk = -1;
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int chunk, b, th, tw;
        dw_tile_coords(p, tile, chunk, b, th, tw);
        const int c = chunk * CC + CPT * cp;
        const bool c_ok = c < p.C;
        if (chunk != cur_chunk) {
            cur_chunk = chunk;
#pragma unroll
            for (int q = 0; q < K * K; ++q) {
#pragma unroll
                for (int e = 0; e < CPT; ++e) wr[q][e] = c_ok ? __ldg(p.w + (size_t)q * p.C + c + e) : 0.f;
            }
            if (c_ok) {
#pragma unroll
                for (int e = 0; e < CPT; ++e) { sc[e] = __ldg(p.scale + c + e); bi[e] = __ldg(p.bias + c + e); }
            }
        }
        const int ho = th * p.Ht + r;
        const int wo_base = tw * p.Ws + wl0;
        const int n_out = max(0, min(p.n_t, p.Wo - wo_base));
        const bool row_ok = ho < p.Ho;
        float* orow = p.out + ((size_t)(b * p.Ho + ho) * p.Wo + wo_base) * p.C + c;
        float sum[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) sum[e] = 0.f;

        mbar_wait(bar_full(stage), phase);
        // tile[row][col][CC ch]: this thread reads rows r*S + kh, columns wl0*S + j.  32-bit shared addresses, one per
        // kernel row, advanced once per group of G columns; the columns inside a group are immediate offsets.
        uint32_t rowaddr[K];
        {
            const uint32_t a0 = base + (uint32_t)stage * (uint32_t)p.tile_bytes +
                                (uint32_t)(((r * S) * p.Wbox + wl0 * S) * (CC * 4) + cp * (4 * CPT));
            const uint32_t rstride = (uint32_t)(p.Wbox * CC * 4);
#pragma unroll
            for (int kh = 0; kh < K; ++kh) rowaddr[kh] = a0 + kh * rstride;
        }
        if (row_ok && c_ok && n_out > 0) {
            float acc[K][CPT];
            const int n_cols = (n_out - 1) * S + K;
            constexpr int G = K * S;
            // one input column (relative index jb + JJ): load its K rows, add it into the outputs it touches, finish
            // the output whose last tap it is.  FIRST: outputs with negative index exist (skip their store);
            // TAIL: the column itself may lie beyond the segment.
            auto column = [&](int jb, int jj, bool first, bool tail) {
                const int j = jb + jj;
                if (tail && j >= n_cols) return;
                float x[K][CPT];
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    if constexpr (CPT == 2)
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                                     : "=f"(x[kh][0]), "=f"(x[kh][1]) : "r"(rowaddr[kh] + (uint32_t)(jj * CC * 4)));
                    else
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[kh][0]) : "r"(rowaddr[kh] + (uint32_t)(jj * CC * 4)));
                }
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                    if ((jj - kw) % S != 0) continue;
                    const int slot = ((((jj - kw) / S) % K) + K) % K;
                    float t[CPT];
#pragma unroll
                    for (int e = 0; e < CPT; ++e) {
                        t[e] = kw == 0 ? x[0][e] * wr[0][e] : fmaf(x[0][e], wr[kw][e], acc[slot][e]);
#pragma unroll
                        for (int kh = 1; kh < K; ++kh) t[e] = fmaf(x[kh][e], wr[kh * K + kw][e], t[e]);
                        acc[slot][e] = t[e];
                    }
                    if (kw == K - 1) {                     // last tap: output u = (j - kw) / S is complete
                        if (!first || j >= kw) {
                            float o[CPT];
#pragma unroll
                            for (int e = 0; e < CPT; ++e) { o[e] = fast_swish(fmaf(t[e], sc[e], bi[e])); sum[e] += o[e]; }
                            float* dst = orow + (size_t)((j - kw) / S) * p.C;
                            if constexpr (CPT == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
                            else *dst = o[0];
                        }
                    }
                }
            };
            auto advance = [&]() {
#pragma unroll
                for (int kh = 0; kh < K; ++kh) rowaddr[kh] += (uint32_t)(G * CC * 4);
            };
            int jb = 0;
            if (n_cols >= G) {                             // first group: no column is beyond the segment
#pragma unroll
                for (int jj = 0; jj < G; ++jj) column(0, jj, true, false);
                advance(); jb = G;
                for (; jb + G <= n_cols; jb += G) {        // interior groups: no checks at all
#pragma unroll
                    for (int jj = 0; jj < G; ++jj) column(jb, jj, false, false);
                    advance();
                }
            }
            if (jb < n_cols) {                             // last, partial group
#pragma unroll
                for (int jj = 0; jj < G; ++jj) column(jb, jj, jb == 0, true);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(stage));      // this warp is done reading the stage
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }

        // channel sums of this thread's outputs.  Lanes of a warp that hold the same channels (32 / TPG groups) are
        // combined by shuffles in a fixed order, then one slot per (tile, warp) is written -- no CTA-wide barrier; the
        // SE kernel adds the slots in a fixed order (deterministic).  When a pixel chunk spans several warps (TPG > 32)
        // those warps hold disjoint channels of the SAME slot.
#pragma unroll
        for (int o = 16; o >= TPG; o >>= 1) {
#pragma unroll
            for (int e = 0; e < CPT; ++e) sum[e] += __shfl_xor_sync(0xffffffffu, sum[e], o);
        }
        constexpr int WPS = TPG > 32 ? TPG / 32 : 1;       // warps per slot
        if (c_ok && (TPG >= 32 || lane < TPG)) {
            float* dst = p.partial + (((size_t)b * (p.tiles_h * p.tiles_w) + th * p.tiles_w + tw) * dw_slots(K, CC) + warp / WPS) * p.C + c;
#pragma unroll
            for (int e = 0; e < CPT; ++e) dst[e] = sum[e];
        }
    }
}

static int dw_chunk(int C, int k) { (void)k; return C >= 256 ? 64 : (C >= 32 ? 32 : 16); }

static void dw_tile_shape(int Ho, int Wo, int C, int k, int s, int& CC, int& groups, int& Ht, int& Ws, int& nsub, int& n_t) {
    CC = dw_chunk(C, k);
    groups = DW_GROUP_THREADS / (CC / 2);
    Ht = Ho >= 8 ? 8 : (Ho >= 4 ? 4 : (Ho >= 2 ? 2 : 1));
    nsub = groups / Ht;
    // output columns per thread: 16 when the row is long enough, never below 4
    n_t = 16;
    while (n_t > 4 && nsub * (n_t / 2) >= Wo) n_t /= 2;
    if (s == 2 && n_t > 8) n_t = 8;       // keep the stride-2 input box comparable
    Ws = nsub * n_t;
    // at least two tiles (input box with halo) must fit in shared memory
    const size_t cap = k == 3 ? 110 * 1024 : DW_SMEM_LIMIT;      // 3x3 kernels run 2 CTAs per SM
    while (n_t > 1 && (size_t)((Ht - 1) * s + k) * ((Ws - 1) * s + k) * CC * 4 * 2 > cap - 768) {
        n_t /= 2;
        Ws = nsub * n_t;
    }
}

int dwconv_tiles_per_clip(int Ho, int Wo, int C, int k, int s) {
    int CC, groups, Ht, Ws, nsub, n_t;
    dw_tile_shape(Ho, Wo, C, k, s, CC, groups, Ht, Ws, nsub, n_t);
    return cdiv(Ho, Ht) * cdiv(Wo, Ws) * dw_slots(k, CC);
}

int dwconv_tma(const DwArgs& a, cudaStream_t st) {
    AC_REQUIRE((a.k == 3 || a.k == 5) && (a.s == 1 || a.s == 2), "dwconv_tma: unsupported k=%d s=%d", a.k, a.s);
    AC_REQUIRE(a.C % 4 == 0 && ((uintptr_t)a.in & 15) == 0, "dwconv_tma: C %% 4 and 16-byte aligned input required");
    if (a.B == 0) return AC_OK;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(tensor_map_encode_fn());
    AC_REQUIRE(encode != nullptr, "dwconv_tma: cuTensorMapEncodeTiled is not available from the driver");
    DwParams p;
    p.out = a.out; p.partial = a.partial; p.w = a.w; p.scale = a.scale; p.bias = a.bias;
    p.B = a.B; p.Ho = a.Ho; p.Wo = a.Wo; p.C = a.C; p.pad_lo = a.pad_lo;
    dw_tile_shape(a.Ho, a.Wo, a.C, a.k, a.s, p.CC, p.groups, p.Ht, p.Ws, p.nsub, p.n_t);
    p.tiles_h = cdiv(a.Ho, p.Ht); p.tiles_w = cdiv(a.Wo, p.Ws); p.chunks = cdiv(a.C, p.CC);
    p.total_tiles = p.chunks * a.B * p.tiles_h * p.tiles_w;
    p.Hbox = (p.Ht - 1) * a.s + a.k; p.Wbox = (p.Ws - 1) * a.s + a.k;
    AC_REQUIRE(p.Wbox <= 256 && p.Hbox <= 256, "dwconv_tma: tile too large");
    p.tile_bytes = (int)align_up((size_t)p.Hbox * p.Wbox * p.CC * 4, 128);
    const int fixed = 128 + 128;                  // alignment slack + barriers
    const int ctas_per_sm = a.k == 3 ? 2 : 1;     // the 5x5 kernels need > 113 registers (measured slower when capped)
    const int smem_cap = a.k == 3 ? 110 * 1024 : DW_SMEM_LIMIT;
    p.stages = std::min(DW_MAX_STAGES, (smem_cap - fixed) / p.tile_bytes);
    AC_REQUIRE(p.stages >= 2, "dwconv_tma: tile of %d bytes does not fit twice in shared memory", p.tile_bytes);
    const size_t smem = (size_t)p.stages * p.tile_bytes + fixed;

    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)a.Wi * a.C * 4, (cuuint64_t)a.Hi * a.Wi * a.C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)p.CC, (cuuint32_t)p.Wbox, (cuuint32_t)p.Hbox, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AC_REQUIRE(cr == CUDA_SUCCESS, "dwconv_tma: cuTensorMapEncodeTiled failed (%d) C=%d Wi=%d Hi=%d B=%d box %dx%d", (int)cr,
               a.C, a.Wi, a.Hi, a.B, p.Wbox, p.Hbox);
    const int grid = std::min(p.total_tiles, kNumSMs * ctas_per_sm);
    AC_TIMED(a.k == 3 ? "dwconv_k3" : "dwconv_k5", st);
#define AC_DW_TMA(K, S, CC)                                                                                      \
    do {                                                                                                         \
        static cudaError_t attr_rc = cudaFuncSetAttribute(dwconv_tma_kernel<K, S, CC>,                           \
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_LIMIT); \
        AC_CUDA(attr_rc);                                                                                        \
        cudaLaunchConfig_t cfg = {};                                                                             \
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(dw_compute_threads(K) + 32); cfg.dynamicSmemBytes = smem; cfg.stream = st; \
        cudaLaunchAttribute at[1] = {pdl_attr()};                                                                \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                        \
        AC_CUDA(cudaLaunchKernelEx(&cfg, dwconv_tma_kernel<K, S, CC>, map, p));                                  \
    } while (0)
#define AC_DW_KS(CC)                                         \
    do {                                                     \
        if (a.k == 3 && a.s == 1) AC_DW_TMA(3, 1, CC);       \
        else if (a.k == 3 && a.s == 2) AC_DW_TMA(3, 2, CC);  \
        else if (a.k == 5 && a.s == 1) AC_DW_TMA(5, 1, CC);  \
        else AC_DW_TMA(5, 2, CC);                            \
    } while (0)
    if (p.CC == 64) AC_DW_KS(64);
    else if (p.CC == 32) AC_DW_KS(32);
    else AC_DW_KS(16);
#undef AC_DW_KS
#undef AC_DW_TMA
    AC_LAUNCHED("dwconv_tma_kernel");
    return AC_OK;
}

}  // namespace ac

// Diagnostic entry point (tests): one depthwise layer on caller buffers; w_dev is [k*k][C] (tap-major).
extern "C" int ac_dwconv(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev,
                         float* out_dev, float* partial_dev, int B, int Hi, int Wi, int C, int k, int s, int pad_lo,
                         int pad_hi, void* stream) {
    using namespace ac;
    AC_REQUIRE(in_dev && w_dev && scale_dev && bias_dev && out_dev && partial_dev, "ac_dwconv: null argument");
    DwArgs a;
    a.in = in_dev; a.out = out_dev; a.partial = partial_dev; a.w = w_dev; a.scale = scale_dev; a.bias = bias_dev;
    a.B = B; a.Hi = Hi; a.Wi = Wi; a.C = C; a.k = k; a.s = s; a.pad_lo = pad_lo;
    a.Ho = (Hi + pad_lo + pad_hi - k) / s + 1;
    a.Wo = (Wi + pad_lo + pad_hi - k) / s + 1;
    return dwconv_tma(a, (cudaStream_t)stream);
}
extern "C" int ac_dwconv_partial_rows(int Ho, int Wo, int C, int k, int s) { return ac::dwconv_tiles_per_clip(Ho, Wo, C, k, s); }

