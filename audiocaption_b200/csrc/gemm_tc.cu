// Tensor-core pointwise GEMM for sm_100a (see gemm.cuh for the operation):
//     C[m, n] = act( (sum_k A[m,k] * ascale[m / rows_per_group, k] * W[n,k]) * cscale[n] + cbias[n] ) + R[m, n]
//
// Numerics: fp32 in, fp32 out.  Every operand is split on the fly into a tf32 "hi" part and an fp32
// residual "lo" part and the product is accumulated as lo*hi + hi*lo + hi*hi in fp32 (TMEM), the
// classic 3xTF32 scheme: the result carries fp32-level accuracy (error ~2^-22 per product), which is
// what keeps the greedy token ids identical to the fp32 reference.
//
// Structure (one persistent CTA per SM, 20 warps, warp-specialised):
//   warp 0       TMA producer: per 32-wide k-chunk one cp.async.bulk.tensor (A box 128 x 32 fp32, 128B
//                swizzle, rows/cols beyond M/K zero-filled by the TMA unit).  The pre-packed weight
//                (hi and lo images, see TcWeight) is either fetched ONCE per CTA with a single bulk copy
//                ("resident": one n-tile and <= 64 KB -- every large-M layer) or streamed per stage.
//   warps 12-19  transform: thread = one row (TMEM lane) x 16 k: reads the landed chunk from shared memory (the
//                128B swizzle makes the row-per-lane reads conflict-free), a *= SE gate, hi = tf32(a),
//                lo = a - hi, and writes both images straight into TENSOR MEMORY (tcgen05.st) -- the A operand
//                never goes back to shared memory; the gate values of the next chunk are prefetched.
//   warps 1, 2   MMA issuers, alternating k-chunks (warp 2 also allocates the tensor memory): tcgen05.mma kind::tf32
//                with A from TMEM and W from shared memory, M=128, N=BN, K=8 per instruction, 3 per k-step (lo*hi,
//                hi*lo, hi*hi); accumulators double-buffered in TMEM; tcgen05.commit releases the smem stage and the
//                TMEM A slot.  A commit parks its thread while the pipe drains, hence two issuers in ping-pong.
//   warps 4-11   epilogue (two warps per TMEM lane quarter, alternating 32-column panels): tcgen05.ld,
//                folded BN scale/bias, swish/relu, residual, then the panel goes through a 128B-swizzled
//                4 KB shared-memory slab and out with one TMA tensor store (coalesced, clipped at M/N).
//                Overlaps the next tile's main loop.
//   warp 3       gatekeeper: executes the mbarrier waits that gate each k-chunk (weights landed, A slot filled,
//                accumulator free) on behalf of the MMA issuer and publishes a running count in shared memory; the
//                issuer only polls that word -- an mbarrier.try_wait issued behind tcgen05.commit costs it 150-350
//                cycles even when the phase completed long ago ("The MMA issuer's instruction stream").
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace ac {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                          // fp32 elements per k-chunk = one 128-byte swizzle row
constexpr int TC_MAX_BN_RESIDENT = 160;
constexpr int TC_MAX_BN_STREAM = 128;
constexpr int TC_RESIDENT_W_BYTES = 64 * 1024;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_XF_WARPS = 8;
constexpr int TC_FIRST_EPI_WARP = 4;
constexpr int TC_FIRST_XF_WARP = TC_FIRST_EPI_WARP + TC_EPI_WARPS;  // 12
constexpr int TC_THREADS = (TC_FIRST_XF_WARP + TC_XF_WARPS) * 32;   // 640
constexpr int TC_A_TILE_BYTES = TC_BM * TC_BK * 4; // 16 KB (hi image; lo image follows)
constexpr int TC_SLAB_BYTES = 32 * 32 * 4;         // one epilogue warp's 32 rows x 32 columns
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_BAR_BYTES = 320;
constexpr int TC_CONV_SEG_CHUNKS = 16;              // conv: k-chunks per TMEM accumulation segment (K = 512)
constexpr int TC_MAX_ACC = 4;                     // accumulator stages in TMEM (2 for wide tiles, 4 for BN <= 64)

struct TcParams {
    const float* wpacked; float* C; const float* R;
    const float* ascale; const float* cbias;
    int M, N, K, ldc, rows_per_group;
    int BN, n_tiles, m_tiles, k_chunks, stages, tmem_cols, resident, a_slots, acc_stages;
    // 3x3 convolution mode (CONV): A is the NHWC activation [cB, cH, cW, Cin] seen through a 4-D tensor map; an
    // m-tile is cBbox clips x cHbox rows x cW columns (<= 128 pixels), k-chunk kc = (tap, 32-channel chunk)
    int cW, cH, cB, cHbox, cBbox, c_tiles_h, c_cpc, a_bytes;
    // CONV: the k loop is cut into segments of seg_chunks chunks; each segment accumulates in TMEM from zero and the
    // epilogue warps add the segments in fp32 (round-to-nearest).  The tensor core truncates when it accumulates, a
    // bias that grows linearly with the number of MMAs: ~2e-4 relative over K = 18432, ~5e-6 over a 512-wide segment.
    int seg_chunks;
    // merge != 0: the TMEM A ring has as many slots as there are smem stages, slot index == stage index, and ONE
    // tcgen05.commit per chunk (on bar_empty) releases both (every commit costs the issuing thread ~165 cycles)
    int merge;
    // passes = 3: every fp32 product is issued as three TF32 MMAs (lo*hi + hi*lo + hi*hi, fp32-level accuracy);
    // passes = 1: the hi*hi MMA only (plain TF32, ~2^-11 relative per product): the "tf32" precision mode of the
    // Cnn14 / SED convolutions for the bf16-class configurations (BASELINE configs[2..4]).  The lo halves are then
    // neither fetched (streamed weights) nor written to tensor memory.
    int passes;
    long long* dbg;   // optional pipeline trace of CTA 0 (AC_TC_TRACE): [event kind 0..7][256] clock64 stamps
};

#define AC_TC_STAMP(kind, idx)                                                               \
    do {                                                                                     \
        if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 256) p.dbg[(kind) * 256 + (idx)] = clock64(); \
    } while (0)

template <int ACT>
__device__ __forceinline__ float4 epi_math(const uint32_t* u, float4 b) {
    float v[4] = {__uint_as_float(u[0]) + b.x, __uint_as_float(u[1]) + b.y, __uint_as_float(u[2]) + b.z,
                  __uint_as_float(u[3]) + b.w};
    if (ACT == ACT_SWISH) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fast_swish(v[e]);
    } else if (ACT == ACT_RELU) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.0f);
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

// The BN scale is folded into the packed weight, so the epilogue is act(acc + bias) + R.
template <int ACT, bool GATED, bool RESID, bool CONV = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapC, const TcParams p) {
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;          // 128B swizzle needs 1024-byte aligned tiles
    // smem: [epilogue slabs][resident W][stages: raw A chunk | (streamed W hi | lo)][barriers][per-warp bias]
    // TMEM: [acc_stages accumulators (BN columns each)][A ring: a_slots x (hi 32 columns | lo 32 columns)]
    const uint32_t slabs = base;
    const uint32_t w_res = slabs + (CONV ? (uint32_t)(TC_BM * p.BN * 4) : (uint32_t)(TC_EPI_WARPS * TC_SLAB_BYTES));   // CONV: fp32 running sums [BN][128]
    const uint32_t w_chunk_bytes = (uint32_t)p.BN * 256u;
    const uint32_t w_load_bytes = p.passes == 1 ? (uint32_t)p.BN * 128u : w_chunk_bytes;   // hi image only in tf32 mode
    const uint32_t stage0 = w_res + (p.resident ? (uint32_t)p.k_chunks * w_chunk_bytes : 0u);
    const uint32_t stage_bytes = TC_A_TILE_BYTES + (p.resident ? 0u : w_load_bytes);
    const uint32_t bars = stage0 + p.stages * stage_bytes;
    auto bar_tma = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 64u + 8u * s; };
    auto bar_axf = [&](int a) { return bars + 128u + 8u * a; };       // A slot (TMEM) holds hi / lo
    auto bar_aempty = [&](int a) { return bars + 160u + 8u * a; };    // A slot consumed by the tensor core
    auto bar_acc_full = [&](int a) { return bars + 192u + 8u * a; };
    auto bar_acc_empty = [&](int a) { return bars + 224u + 8u * a; };
    const uint32_t bar_w = bars + 256u;
    const uint32_t tmem_slot_addr = bars + 264u;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot_addr - raw));
    volatile int* s_ready = reinterpret_cast<volatile int*>(smem_raw + (tmem_slot_addr + 4u - raw));   // chunks cleared for the MMA issuer
    float* bias_all = reinterpret_cast<float*>(smem_raw + (bars + TC_BAR_BYTES - raw));   // [EPI_WARPS][TC_MAX_BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    if (threadIdx.x == 0) AC_TC_STAMP(7, 255);           // kernel entry

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapA);
        prefetch_tensormap(&mapC);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_tma(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < p.a_slots; ++a) {
            mbar_init(bar_axf(a), TC_XF_WARPS);
            mbar_init(bar_aempty(a), 1);
        }
        for (int a = 0; a < p.acc_stages; ++a) {
            mbar_init(bar_acc_full(a), 1);
            mbar_init(bar_acc_empty(a), TC_EPI_WARPS);
        }
        mbar_init(bar_w, 1);
        *s_ready = 0;
        s_ready[1] = 0;                                  // "MMAs issued" chunk count (issuer ping-pong)
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc(tmem_slot_addr, (uint32_t)p.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) AC_TC_STAMP(7, 254);           // barriers initialised, tensor memory allocated
    pdl_trigger();                                   // the next kernel may start its own prologue
    if (warp == 0 && lane == 0 && p.resident) {      // weights are constants: fetch them before the dependency wait
        const uint32_t wb = (uint32_t)p.k_chunks * w_chunk_bytes;
        mbar_expect_tx(bar_w, wb);
        bulk_load(w_res, p.wpacked, wb, bar_w);
    }
    pdl_wait();                                      // A, the SE gate, the residual: produced by earlier kernels
    if (threadIdx.x == 0) AC_TC_STAMP(7, 253);           // dependencies resolved

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int ev = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_t = tile % p.n_tiles, m_t = tile / p.n_tiles;
                const float* wsrc = p.wpacked + (size_t)n_t * p.k_chunks * (p.BN * 64);
                int tap = 0, cc = 0;                       // CONV: kc = tap * c_cpc + cc
                const int c_b0 = CONV ? (m_t / p.c_tiles_h) * p.cBbox : 0;
                const int c_h0 = CONV ? (m_t % p.c_tiles_h) * p.cHbox : 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(bar_empty(stage), phase ^ 1u);
                    AC_TC_STAMP(0, ev); ++ev;
                    const uint32_t sa = stage0 + stage * stage_bytes;
                    mbar_expect_tx(bar_tma(stage), (uint32_t)p.a_bytes + (p.resident ? 0u : w_load_bytes));
                    if (CONV) {
                        // the tile's pixels shifted by the tap; rows / columns outside the image arrive as zeros
                        tma_load_4d(sa, &mapA, cc * TC_BK, tap % 3 - 1, c_h0 + tap / 3 - 1, c_b0, bar_tma(stage));
                        if (++cc == p.c_cpc) { cc = 0; ++tap; }
                    } else {
                        tma_load_2d(sa, &mapA, kc * TC_BK, m_t * TC_BM, bar_tma(stage));
                    }
                    if (!p.resident)
                        bulk_load(sa + TC_A_TILE_BYTES, wsrc + (size_t)kc * (p.BN * 64), w_load_bytes, bar_tma(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ gatekeeper: does the MMA issuer's barrier waits
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            int n = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int in_seg = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    if (in_seg == 0) mbar_wait(bar_acc_empty(acc), acc_phase ^ 1u);   // the segment's accumulator is free
                    if (!p.resident) mbar_wait(bar_tma(stage), phase);
                    mbar_wait(bar_axf(as), aphase);
                    ++n;
                    ++in_seg;
                    if (kc + 1 == p.k_chunks || (CONV && in_seg == p.seg_chunks)) {
                        in_seg = 0;
                        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
                    }
                    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(tmem_slot_addr + 4u), "r"(n) : "memory");
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------ MMA issuers (two warps, alternating k-chunks)
        // tcgen05.commit parks the thread that executes it for ~330 cycles while the tensor pipe drains, so ONE issuer
        // leaves the pipe idle between chunks.  Two issuers ping-pong: while one sits in its commit the other (which has
        // seen the gatekeeper's clearance and the first one's "MMAs issued" flag) is already feeding the next chunk.
        // MMAs reach the pipe in flag order, and the pipe retires them in order, so a commit by either thread also
        // covers the other thread's earlier chunks.
        const int me = warp - 1;
        // The whole warp runs this loop converged (all operands are warp-uniform, so descriptors live in uniform
        // registers and nothing is recomputed per lane); one elected lane issues the tcgen05 instructions.  The MMA
        // thread's own instruction stream is the limiter for small N, so the per-instruction work is kept minimal:
        // descriptors are formed once per stage and advanced by adding to their low word.
        const uint32_t idesc = umma_idesc(2, TC_BM, p.BN);
        const uint32_t a_ring = tmem_base + (uint32_t)(p.acc_stages * p.BN);
        const uint64_t desc_hi_bits = umma_desc_sw128(0) & 0xffffffff00000000ull;   // SBO / version / swizzle fields
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int n_chunk = 0;                                  // running chunk index of this CTA (both issuers count all)
        if (p.resident) mbar_wait(bar_w, 0u);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            uint32_t d_tmem = 0;
            int in_seg = 0;                                   // chunk index within the current accumulation segment
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                if (in_seg == 0) d_tmem = tmem_base + (uint32_t)(acc * p.BN);   // (the gatekeeper checked that it is free)
                const bool mine = (n_chunk & 1) == me;
                if (mine) {
                    // chunk cleared by the gatekeeper, previous chunk's MMAs issued by the other warp (plain shared-memory
                    // words: no mbarrier instruction on the issuing threads)
                    int seen;
                    do {
                        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(seen) : "r"(tmem_slot_addr + 4u) : "memory");
                    } while (seen <= n_chunk);
                    do {
                        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(seen) : "r"(tmem_slot_addr + 8u) : "memory");
                    } while (seen < n_chunk);
                    tc_fence_after();
                }
                if (mine && lane == 0) AC_TC_STAMP(3, n_chunk);
                const uint32_t sa = stage0 + stage * stage_bytes;
                const uint32_t a_hi = a_ring + (uint32_t)as * 64u, a_lo = a_hi + 32u;
                const uint32_t w_hi = p.resident ? w_res + kc * w_chunk_bytes : sa + TC_A_TILE_BYTES;
                const uint32_t w_lo = w_hi + (uint32_t)p.BN * 128u;
                const uint64_t dwh0 = desc_hi_bits | (uint64_t)((w_hi & 0x3ffffu) >> 4);
                const uint64_t dwl0 = desc_hi_bits | (uint64_t)((w_lo & 0x3ffffu) >> 4);
                const int ksteps = min(TC_BK / 8, (p.K - kc * TC_BK) / 8);
                const bool seg_end = kc + 1 == p.k_chunks || (CONV && in_seg + 1 == p.seg_chunks);
                if (leader && mine) {
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        if (ks < ksteps) {
                            // 8 tf32 = 32 bytes along the swizzled W row = +2 in the descriptor's address field
                            if (p.passes == 1) {
                                mma_tf32_ts(d_tmem, a_hi + ks * 8u, dwh0 + 2u * ks, idesc, (in_seg | ks) != 0 ? 1u : 0u);
                            } else {
                                mma_tf32_ts(d_tmem, a_lo + ks * 8u, dwh0 + 2u * ks, idesc, (in_seg | ks) != 0 ? 1u : 0u);   // small terms first
                                mma_tf32_ts(d_tmem, a_hi + ks * 8u, dwl0 + 2u * ks, idesc, 1u);
                                mma_tf32_ts(d_tmem, a_hi + ks * 8u, dwh0 + 2u * ks, idesc, 1u);
                            }
                        }
                    }
                    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(tmem_slot_addr + 8u), "r"(n_chunk + 1) : "memory");
                    AC_TC_STAMP(7, n_chunk);
                    mma_commit(bar_empty(stage));         // smem stage reusable once these MMAs retire
                    if (!p.merge) mma_commit(bar_aempty(as));   // ... and so is the TMEM A slot
                    if (seg_end) mma_commit(bar_acc_full(acc));
                }
                __syncwarp();
                if (mine && lane == 0) { AC_TC_STAMP(4, n_chunk); }
                ++n_chunk;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
                ++in_seg;
                if (seg_end) {
                    in_seg = 0;
                    if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else if (warp >= TC_FIRST_XF_WARP) {
        // ------------------------------------------------------------ transform (gate, hi/lo split -> TMEM)
        const int wq = warp & 3;                                      // TMEM lane quarter of this warp
        const int wh = (warp - TC_FIRST_XF_WARP) >> 2;                // which 16 of the chunk's 32 k
        const int r = wq * 32 + lane;                                 // row of the tile = TMEM lane
        const uint32_t a_ring = tmem_base + (uint32_t)(p.acc_stages * p.BN) + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wh * 16);
        const float* grow = nullptr;                                  // gate row of this thread's row
        auto gate_row = [&](int tile) {
            const int row = min((tile / p.n_tiles) * TC_BM + r, p.M - 1);
            grow = p.ascale + (size_t)(row / p.rows_per_group) * p.K + wh * 16;
        };
        auto load_gate = [&](int kc, float4 (&g)[4]) {
            const int k = kc * TC_BK + wh * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                g[j] = k + 4 * j < p.K ? __ldg(reinterpret_cast<const float4*>(grow + kc * TC_BK + 4 * j))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        int tile = blockIdx.x, kc = 0;
        int xev = 0;
        const bool t0 = threadIdx.x == TC_FIRST_XF_WARP * 32;
        float4 g_next[4];
        if (GATED && tile < total_tiles) { gate_row(tile); load_gate(kc, g_next); }
        while (tile < total_tiles) {
            float4 g[4];
            if (GATED) {
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] = g_next[j];
            }
            int kc2 = kc + 1, tile2 = tile;
            if (kc2 == p.k_chunks) { kc2 = 0; tile2 += gridDim.x; }
            if (GATED && tile2 < total_tiles) {                       // next chunk's gates in flight meanwhile
                if (kc2 == 0) gate_row(tile2);
                load_gate(kc2, g_next);
            }
            mbar_wait(bar_tma(stage), phase);
            if (t0) AC_TC_STAMP(1, xev);
            // row r of the 128B-swizzled chunk: logical 16-byte piece c sits at piece c ^ (r & 7)
            const uint8_t* arow = smem_raw + (stage0 - raw) + (size_t)stage * stage_bytes + r * 128;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(arow + (((wh * 4 + j) ^ (r & 7)) << 4));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                if (GATED) { x[0] *= g[j].x; x[1] *= g[j].y; x[2] *= g[j].z; x[3] *= g[j].w; }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = tf32_rna(x[e]);
                    hi[4 * j + e] = __float_as_uint(h);
                    lo[4 * j + e] = __float_as_uint(x[e] - h);
                }
            }
            if (p.merge) mbar_wait(bar_empty(stage), phase ^ 1u);   // (slot == stage: released together)
            else mbar_wait(bar_aempty(as), aphase ^ 1u);   // the tensor core is done with this TMEM A slot
            tc_fence_after();
            tmem_st16(a_ring + (uint32_t)as * 64u, hi);
            if (p.passes != 1) tmem_st16(a_ring + (uint32_t)as * 64u + 32u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_axf(as));
            if (t0) AC_TC_STAMP(2, xev);
            ++xev;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
            tile = tile2; kc = kc2;
        }
    } else if (warp >= TC_FIRST_EPI_WARP) {
        // ------------------------------------------------------------ epilogue
        const int ew = warp - TC_FIRST_EPI_WARP;
        const int q = warp & 3;                            // TMEM lane quarter this warp may read
        const int half = ew >> 2;                          // which of the two warps of that quarter
        const uint32_t slab = slabs + ew * TC_SLAB_BYTES;
        uint8_t* slab_gen = smem_raw + (slab - raw);
        float* bias = bias_all + ew * TC_MAX_BN_RESIDENT;  // this warp's private copy (zero beyond N)
        const int ldc = p.ldc;
        const int full_panels = p.BN / 32;
        const bool tail16 = (p.BN & 31) != 0;
        int acc = 0; uint32_t acc_phase = 0;
        int cur_nt = -1;
        int eev = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_t = tile % p.n_tiles, m_t = tile / p.n_tiles;
            const int row0 = m_t * TC_BM + q * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            const float* rrow = p.R + (size_t)(row_ok ? row : 0) * ldc;
            if (n_t != cur_nt) {
                cur_nt = n_t;
                __syncwarp();
                for (int c = lane; c < p.BN; c += 32) {
                    const int col = n_t * p.BN + c;
                    bias[c] = (p.cbias != nullptr && col < p.N) ? __ldg(p.cbias + col) : 0.0f;
                }
                __syncwarp();
            }
            mbar_wait(bar_acc_full(acc), acc_phase);
            tc_fence_after();
            if (ew == 0 && lane == 0) AC_TC_STAMP(5, eev);
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            if (CONV) {
                // row of the tile -> pixel (clip, h, w); every thread stores its pixel's 32-channel runs directly
                const int r = q * 32 + lane;
                const int w_ = r % p.cW, t_ = r / p.cW;
                const int hl = t_ % p.cHbox, bl = t_ / p.cHbox;
                const int h_ = (m_t % p.c_tiles_h) * p.cHbox + hl, b_ = (m_t / p.c_tiles_h) * p.cBbox + bl;
                const bool ok = bl < p.cBbox && h_ < p.cH && b_ < p.cB;
                float* crow = p.C + (((size_t)b_ * p.cH + h_) * p.cW + w_) * ldc;
                float* sums = reinterpret_cast<float*>(slab_gen - ew * TC_SLAB_BYTES) + r;   // running sums [BN][128], mine: [.][r]
                const int nseg = (p.k_chunks + p.seg_chunks - 1) / p.seg_chunks;
                for (int sg = 0; sg < nseg; ++sg) {
                    if (sg != 0) { mbar_wait(bar_acc_full(acc), acc_phase); tc_fence_after(); }
                    const bool first = sg == 0, last = sg + 1 == nseg;
                    const uint32_t t_seg = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
                    const int my_items = full_panels + (tail16 ? 1 : 0);
                    for (int pn = half; pn < my_items; pn += 2) {
                        const bool is_tail = pn == full_panels;          // trailing 16-column half panel
                        const int col0 = n_t * p.BN + pn * 32;
                        uint32_t u[2][16];
                        tmem_ld16(t_seg + (uint32_t)(pn * 32), u[0]);
                        if (!is_tail) tmem_ld16(t_seg + (uint32_t)(pn * 32 + 16), u[1]);
                        tmem_ld_wait();
                        float* sp = sums + (size_t)(pn * 32) * TC_BM;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (c < 16 || !is_tail) {
                                float v = __uint_as_float(u[c >> 4][c & 15]);
                                if (!first) v += sp[c * TC_BM];
                                if (!last) sp[c * TC_BM] = v;
                                u[c >> 4][c & 15] = __float_as_uint(v);
                            }
                        }
                        if (last) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 bq = *reinterpret_cast<const float4*>(bias + pn * 32 + 4 * j);
                                const float4 o = epi_math<ACT>(&u[j >> 2][(j & 3) * 4], bq);
                                if (ok && (j < 4 || !is_tail) && col0 + 4 * j < p.N)
                                    *reinterpret_cast<float4*>(crow + col0 + 4 * j) = o;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty(acc));
                    if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
                }
                if (ew == 0 && lane == 0) AC_TC_STAMP(6, eev);
                ++eev;
                continue;
            }
            // a tile with a single work item (one panel or only the half panel) alternates between the two warps of
            // a lane quarter from tile to tile, so two tiles' epilogues overlap
            const bool single = full_panels + (tail16 ? 1 : 0) == 1;
            const int first_pn = single ? ((eev & 1) == half ? 0 : full_panels) : half;
            const bool tail_mine = tail16 && (single ? (eev & 1) == half : half == (full_panels & 1));
            for (int pn = first_pn; pn < full_panels; pn += 2) {
                const int col0 = n_t * p.BN + pn * 32;
                float4 r[8];
                if (RESID) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        r[j] = (row_ok && col0 + 4 * j < p.N) ? __ldg(reinterpret_cast<const float4*>(rrow + col0 + 4 * j))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint32_t u[2][16];
                tmem_ld16(t_row + (uint32_t)(pn * 32), u[0]);
                tmem_ld16(t_row + (uint32_t)(pn * 32 + 16), u[1]);
                tmem_ld_wait();
                float4 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + pn * 32 + 4 * j);
                    o[j] = epi_math<ACT>(&u[j >> 2][(j & 3) * 4], b);
                    if (RESID) { o[j].x += r[j].x; o[j].y += r[j].y; o[j].z += r[j].z; o[j].w += r[j].w; }
                }
                if (lane == 0) bulk_wait_read0();          // previous panel's store has drained the slab
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(slab_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) = o[j];
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&mapC, slab, col0, row0);
                    bulk_commit();
                }
            }
            if (tail_mine) {
                // trailing 16-column half panel: direct 128-bit stores
                const int col0 = n_t * p.BN + full_panels * 32;
                float4 r[4];
                if (RESID) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        r[j] = (row_ok && col0 + 4 * j < p.N) ? __ldg(reinterpret_cast<const float4*>(rrow + col0 + 4 * j))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint32_t u[16];
                tmem_ld16(t_row + (uint32_t)(full_panels * 32), u);
                tmem_ld_wait();
                float* crow = p.C + (size_t)row * ldc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + full_panels * 32 + 4 * j);
                    float4 o = epi_math<ACT>(&u[4 * j], b);
                    if (RESID) { o.x += r[j].x; o.y += r[j].y; o.z += r[j].z; o.w += r[j].w; }
                    if (row_ok && col0 + 4 * j < p.N) *reinterpret_cast<float4*>(crow + col0 + 4 * j) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(acc));
            if (ew == 0 && lane == 0) AC_TC_STAMP(6, eev);
            ++eev;
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) bulk_wait_read0();
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) AC_TC_STAMP(7, 252);           // all roles done
    if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static void tc_tiling(int N, int K, int& BN, int& n_tiles, int& k_chunks, int bn_cap = 0);

// ------------------------------------------------------------------------------------ weight packing
// One thread per packed float4: [n_tile][k_chunk][hi|lo][row BN][16-byte chunk 8 (swizzled)].
__global__ void tc_pack_kernel(const float* __restrict__ W, const float* __restrict__ scale, float* __restrict__ dst,
                               int N, int K, int BN, int n_tiles, int k_chunks, int64_t sn, int64_t sk) {
    const int64_t total = (int64_t)n_tiles * k_chunks * 2 * BN * 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int pc = (int)(i % 8);
    int64_t r = i / 8;
    const int row = (int)(r % BN); r /= BN;
    const int hl = (int)(r % 2); r /= 2;
    const int kc = (int)(r % k_chunks);
    const int n_t = (int)(r / k_chunks);
    const int lc = pc ^ (row & 7);
    const int n = n_t * BN + row, k = kc * TC_BK + lc * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        const float sc = scale != nullptr ? scale[n] : 1.0f;
        for (int e = 0; e < 4; ++e)
            if (k + e < K) v[e] = W[(size_t)n * sn + (size_t)(k + e) * sk] * sc;
    }
    float o[4];
    for (int e = 0; e < 4; ++e) {
        const float h = ptx::tf32_rna(v[e]);
        o[e] = hl == 0 ? h : v[e] - h;
    }
    reinterpret_cast<float4*>(dst)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// Many matrices in ONE launch (the training path re-packs ~70 weight images per step): job j covers packed float4 items
// [first_j, first_{j+1}); a thread finds its job by binary search in the (device-resident, static) job table.
__global__ void tc_pack_multi_kernel(const TcPackJob* __restrict__ jobs, int n_jobs, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const TcPackJob j = jobs[lo];
    long long r = i - j.first;
    const int pc = (int)(r % 8); r /= 8;
    const int row = (int)(r % j.BN); r /= j.BN;
    const int hl = (int)(r % 2); r /= 2;
    const int kc = (int)(r % j.k_chunks);
    const int n_t = (int)(r / j.k_chunks);
    const int lc = pc ^ (row & 7);
    const int n = n_t * j.BN + row, k = kc * TC_BK + lc * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < j.N) {
        for (int e = 0; e < 4; ++e)
            if (k + e < j.K) v[e] = j.W[(size_t)n * j.sn + (size_t)(k + e) * j.sk];
    }
    float o[4];
    for (int e = 0; e < 4; ++e) {
        const float h = ptx::tf32_rna(v[e]);
        o[e] = hl == 0 ? h : v[e] - h;
    }
    reinterpret_cast<float4*>(j.dst)[i - j.first] = make_float4(o[0], o[1], o[2], o[3]);
}

long long tc_pack_plan(const float* W, int N, int K, long long sn, long long sk, float* dst, long long first, TcPackJob* job,
                       TcWeight* out) {
    int BN, n_tiles, k_chunks;
    tc_tiling(N, K, BN, n_tiles, k_chunks);
    job->W = W; job->dst = dst; job->N = N; job->K = K; job->BN = BN; job->n_tiles = n_tiles; job->k_chunks = k_chunks;
    job->sn = sn; job->sk = sk; job->first = first;
    out->packed = dst; out->scale = nullptr; out->N = N; out->K = K; out->BN = BN; out->n_tiles = n_tiles; out->k_chunks = k_chunks;
    return (long long)n_tiles * k_chunks * 2 * BN * 8;
}

int tc_pack_multi(const TcPackJob* jobs_dev, int n_jobs, long long total_items, cudaStream_t st) {
    if (n_jobs <= 0 || total_items <= 0) return AC_OK;
    tc_pack_multi_kernel<<<(unsigned)cdiv64(total_items, 256), 256, 0, st>>>(jobs_dev, n_jobs, total_items);
    AC_LAUNCHED("tc_pack_multi_kernel");
    return AC_OK;
}

// PyTorch Conv2d weight [Cout, Cin, 3, 3] -> [Cout, 9 * Cin] with k = (ky*3+kx) * Cin + c (tap-major: a 32-wide
// k-chunk is 32 consecutive input channels of one tap, i.e. one shifted NHWC box)
__global__ void conv3x3_permute_kernel(const float* __restrict__ w, float* __restrict__ o, int Cout, int Cin) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Cout * Cin * 9) return;
    const int c = (int)(i % Cin);
    const int tap = (int)((i / Cin) % 9);
    const int n = (int)(i / ((int64_t)Cin * 9));
    o[i] = w[((size_t)n * Cin + c) * 9 + tap];
}
int conv3x3_permute_weight(const float* w_dev, float* out_dev, int Cout, int Cin, cudaStream_t st) {
    const int64_t total = (int64_t)Cout * Cin * 9;
    conv3x3_permute_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(w_dev, out_dev, Cout, Cin);
    AC_LAUNCHED("conv3x3_permute_kernel");
    return AC_OK;
}

static bool tc_resident(int BN, int n_tiles, int k_chunks) {
    return n_tiles == 1 && (int64_t)k_chunks * BN * 256 <= TC_RESIDENT_W_BYTES;
}
static void tc_tiling(int N, int K, int& BN, int& n_tiles, int& k_chunks, int bn_cap) {
    k_chunks = cdiv(K, TC_BK);
    n_tiles = 1;
    BN = cdiv(N, 16) * 16;
    if (bn_cap > 0) {      // a narrow image for small-M launches: more n-tiles (more CTAs), weights streamed
        n_tiles = cdiv(N, std::max(16, bn_cap / 16 * 16));
        BN = cdiv(cdiv(N, n_tiles), 16) * 16;
        return;
    }
    if (N <= TC_MAX_BN_RESIDENT && tc_resident(BN, 1, k_chunks)) return;
    static const int bn_stream = [] { const char* e = getenv("AC_TC_BN_STREAM"); const int v = e ? atoi(e) : 0;
                                      return v >= 16 && v <= TC_MAX_BN_STREAM ? v / 16 * 16 : TC_MAX_BN_STREAM; }();   // tuning aid
    n_tiles = cdiv(N, bn_stream);
    BN = cdiv(cdiv(N, n_tiles), 16) * 16;
}

size_t tc_packed_floats(int N, int K) { return tc_packed_floats_bn(N, K, 0); }
size_t tc_packed_floats_bn(int N, int K, int bn_cap) {
    int BN, n_tiles, k_chunks;
    tc_tiling(N, K, BN, n_tiles, k_chunks, bn_cap);
    return (size_t)n_tiles * k_chunks * BN * 64;
}

int tc_pack_weight(const float* W_dev, const float* scale_dev, int N, int K, float* dst_dev, cudaStream_t st,
                   TcWeight* out) {
    return tc_pack_weight_strided(W_dev, scale_dev, N, K, K, 1, dst_dev, st, out);
}

int tc_pack_weight_bn(const float* W_dev, const float* scale_dev, int N, int K, int bn_cap, float* dst_dev, cudaStream_t st,
                      TcWeight* out) {
    AC_REQUIRE(W_dev && dst_dev && out && N > 0 && K > 0, "tc_pack_weight: bad argument");
    AC_REQUIRE(((uintptr_t)dst_dev & 127) == 0, "tc_pack_weight: destination must be 128-byte aligned");
    int BN, n_tiles, k_chunks;
    tc_tiling(N, K, BN, n_tiles, k_chunks, bn_cap);
    const int64_t total = (int64_t)n_tiles * k_chunks * 2 * BN * 8;
    tc_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(W_dev, scale_dev, dst_dev, N, K, BN, n_tiles, k_chunks, K, 1);
    AC_LAUNCHED("tc_pack_kernel");
    out->packed = dst_dev; out->scale = scale_dev; out->N = N; out->K = K; out->BN = BN; out->n_tiles = n_tiles; out->k_chunks = k_chunks;
    return AC_OK;
}

int tc_pack_weight_strided(const float* W_dev, const float* scale_dev, int N, int K, int64_t sn, int64_t sk, float* dst_dev,
                           cudaStream_t st, TcWeight* out) {
    AC_REQUIRE(W_dev && dst_dev && out && N > 0 && K > 0, "tc_pack_weight: bad argument");
    AC_REQUIRE(((uintptr_t)dst_dev & 127) == 0, "tc_pack_weight: destination must be 128-byte aligned");
    int BN, n_tiles, k_chunks;
    tc_tiling(N, K, BN, n_tiles, k_chunks);
    const int64_t total = (int64_t)n_tiles * k_chunks * 2 * BN * 8;
    tc_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(W_dev, scale_dev, dst_dev, N, K, BN, n_tiles, k_chunks, sn, sk);
    AC_LAUNCHED("tc_pack_kernel");
    out->packed = dst_dev; out->scale = scale_dev; out->N = N; out->K = K; out->BN = BN; out->n_tiles = n_tiles; out->k_chunks = k_chunks;
    return AC_OK;
}

// ------------------------------------------------------------------------------------ launch
void* tensor_map_encode_fn() {   // cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
    static void* fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = p;
    });
    return fn;
}

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(tensor_map_encode_fn());
}
static long long* g_tc_trace = nullptr;   // device buffer [8][256], set by ac_gemm_trace

int gemm_tc(const GemmArgs& g, cudaStream_t st) {
    AC_REQUIRE(g.tw != nullptr && g.tw->packed != nullptr, "gemm_tc: weight not packed");
    const TcWeight& w = *g.tw;
    AC_REQUIRE(w.N == g.N && w.K == g.K, "gemm_tc: packed weight is %dx%d, call wants %dx%d", w.N, w.K, g.N, g.K);
    AC_REQUIRE(g.K % 8 == 0 && g.N % 4 == 0, "gemm_tc: K (%d) %% 8 and N (%d) %% 4 must be 0", g.K, g.N);
    AC_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.C & 15) == 0, "gemm_tc: A and C must be 16-byte aligned");
    AC_REQUIRE((g.ldc ? g.ldc : g.N) % 4 == 0, "gemm_tc: ldc must be a multiple of 4");
    if (g.M <= 0 || g.N <= 0) return AC_OK;
    auto encode = tensor_map_encoder();
    AC_REQUIRE(encode != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");

    const int ldc = g.ldc ? g.ldc : g.N;
    CUtensorMap map, mapC;
    {
        const cuuint64_t cdims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
        const cuuint64_t cstr[1] = {(cuuint64_t)ldc * sizeof(float)};
        const cuuint32_t cbox[2] = {32, 32};
        const cuuint32_t ces[2] = {1, 1};
        CUresult cr = encode(&mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g.C, cdims, cstr, cbox, ces,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(cr == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) for C=%p M=%d N=%d ldc=%d", (int)cr,
                   g.C, g.M, g.N, ldc);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.M};
    const int lda = g.lda ? g.lda : g.K;
    AC_REQUIRE(lda % 4 == 0 && lda >= g.K, "gemm_tc: lda (%d) must be a multiple of 4 and >= K (%d)", lda, g.K);
    const cuuint64_t strides[1] = {(cuuint64_t)lda * sizeof(float)};
    const cuuint32_t box[2] = {TC_BK, TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.A), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AC_REQUIRE(cr == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) for A=%p M=%d K=%d", (int)cr, g.A, g.M,
               g.K);

    TcParams p;
    AC_REQUIRE(g.cscale == w.scale, "gemm_tc: the BN scale is folded into the packed weight; cscale must be the packed one");
    p.wpacked = w.packed; p.C = g.C; p.R = g.R; p.ascale = g.ascale; p.cbias = g.cbias;
    p.M = g.M; p.N = g.N; p.K = g.K; p.ldc = ldc; p.rows_per_group = g.rows_per_group > 0 ? g.rows_per_group : 1;
    p.BN = w.BN; p.n_tiles = w.n_tiles; p.m_tiles = cdiv(g.M, TC_BM); p.k_chunks = w.k_chunks;
    p.resident = tc_resident(w.BN, w.n_tiles, w.k_chunks) ? 1 : 0;
    p.passes = g.passes == 1 ? 1 : 3;
    p.dbg = g_tc_trace;
    p.cW = p.cH = p.cB = p.cHbox = p.cBbox = p.c_tiles_h = p.c_cpc = 0; p.a_bytes = TC_A_TILE_BYTES; p.seg_chunks = w.k_chunks;
    const int fixed = 1024 + TC_BAR_BYTES + TC_EPI_WARPS * TC_MAX_BN_RESIDENT * 4 + TC_EPI_WARPS * TC_SLAB_BYTES +
                      (p.resident ? w.k_chunks * w.BN * 256 : 0);
    const int sb = TC_A_TILE_BYTES + (p.resident ? 0 : w.BN * (p.passes == 1 ? 128 : 256));
    p.stages = std::min(TC_MAX_STAGES, (TC_SMEM_LIMIT - fixed) / sb);
    AC_REQUIRE(p.stages >= 2, "gemm_tc: tile too large for shared memory (BN=%d)", w.BN);
    p.tmem_cols = 512;                                         // one CTA per SM owns the whole tensor memory
    p.acc_stages = w.BN <= 64 ? TC_MAX_ACC : 2;
    p.a_slots = std::min(4, (512 - p.acc_stages * w.BN) / 64);
    AC_REQUIRE(p.a_slots >= 2, "gemm_tc: BN=%d leaves no tensor memory for the A operand", w.BN);
    p.merge = p.stages <= p.a_slots ? 1 : 0;
    if (p.merge) p.a_slots = p.stages;
    const size_t smem = (size_t)p.stages * sb + fixed;

    const int grid = std::min(p.m_tiles * p.n_tiles, kNumSMs);
    const bool gated = g.ascale != nullptr, resid = g.R != nullptr;
    AC_TIMED("gemm_tc", st);
#define AC_TC_LAUNCH(ACT, GATED, RESID)                                                                            \
    do {                                                                                                           \
        static cudaError_t attr_rc = cudaFuncSetAttribute(gemm_tc_kernel<ACT, GATED, RESID>,                        \
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT); \
        AC_CUDA(attr_rc);                                                                                          \
        cudaLaunchConfig_t cfg = {};                                                                               \
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;   \
        cudaLaunchAttribute at[1] = {pdl_attr()};                                                                  \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                          \
        AC_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<ACT, GATED, RESID>, map, mapC, p));                        \
    } while (0)
#define AC_TC_ACT(ACT)                                                   \
    do {                                                                 \
        if (gated && resid) AC_TC_LAUNCH(ACT, true, true);               \
        else if (gated) AC_TC_LAUNCH(ACT, true, false);                  \
        else if (resid) AC_TC_LAUNCH(ACT, false, true);                  \
        else AC_TC_LAUNCH(ACT, false, false);                            \
    } while (0)
    if (g.act == ACT_SWISH) AC_TC_ACT(ACT_SWISH);
    else if (g.act == ACT_RELU) AC_TC_ACT(ACT_RELU);
    else AC_TC_ACT(ACT_NONE);
#undef AC_TC_ACT
#undef AC_TC_LAUNCH
    AC_LAUNCHED("gemm_tc_kernel");
    return AC_OK;
}

// 3x3 / stride 1 / pad 1 convolution + folded BN + activation as an implicit GEMM on the same pipeline:
//   out[b,h,w,n] = act( sum_{ky,kx,c} in[b, h+ky-1, w+kx-1, c] * Wp[n, (ky*3+kx)*Cin + c] + bias[n] )
// (captioning/models/cnn_encoder.py:32-75 `ConvBlock`: conv -> BatchNorm -> ReLU, eval mode.)  in/out NHWC fp32;
// the weight is the [Cout, 9*Cin] matrix (tap-major, BN scale folded) packed by tc_pack_weight.
void conv3x3_tile_shape(int H, int W, int& Hbox, int& Bbox) {
    if (W * H <= TC_BM / 2) { Hbox = H; Bbox = TC_BM / (W * H); }
    else { Hbox = std::max(1, std::min(H, TC_BM / W)); Bbox = 1; }
}

int conv3x3_tc(const Conv3Args& a, cudaStream_t st) {
    AC_REQUIRE(a.tw != nullptr && a.tw->packed != nullptr, "conv3x3_tc: weight not packed");
    const TcWeight& w = *a.tw;
    AC_REQUIRE(a.Cin % TC_BK == 0 && a.Cout % 4 == 0, "conv3x3_tc: Cin (%d) %% 32 and Cout (%d) %% 4 must be 0", a.Cin, a.Cout);
    AC_REQUIRE(w.N == a.Cout && w.K == 9 * a.Cin, "conv3x3_tc: packed weight is %dx%d, call wants %dx%d", w.N, w.K,
               a.Cout, 9 * a.Cin);
    AC_REQUIRE(a.W >= 1 && a.W <= TC_BM && a.H >= 1, "conv3x3_tc: width %d must be in 1..128", a.W);
    AC_REQUIRE(((uintptr_t)a.in & 15) == 0 && ((uintptr_t)a.out & 15) == 0, "conv3x3_tc: buffers must be 16-byte aligned");
    if (a.B <= 0) return AC_OK;
    auto encode = tensor_map_encoder();
    AC_REQUIRE(encode != nullptr, "conv3x3_tc: cuTensorMapEncodeTiled is not available from the driver");
    int Hbox, Bbox;
    conv3x3_tile_shape(a.H, a.W, Hbox, Bbox);
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.Cin * 4, (cuuint64_t)a.W * a.Cin * 4, (cuuint64_t)a.H * a.W * a.Cin * 4};
    const cuuint32_t box[4] = {TC_BK, (cuuint32_t)a.W, (cuuint32_t)Hbox, (cuuint32_t)Bbox};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AC_REQUIRE(cr == CUDA_SUCCESS, "conv3x3_tc: cuTensorMapEncodeTiled failed (%d) for in=%p [%d,%d,%d,%d]", (int)cr, a.in,
               a.B, a.H, a.W, a.Cin);

    TcParams p;
    p.wpacked = w.packed; p.C = a.out; p.R = nullptr; p.ascale = nullptr; p.cbias = a.bias;
    p.M = a.B * a.H * a.W; p.N = a.Cout; p.K = 9 * a.Cin; p.ldc = a.Cout; p.rows_per_group = 1;
    p.BN = w.BN; p.n_tiles = w.n_tiles; p.k_chunks = w.k_chunks;
    p.cW = a.W; p.cH = a.H; p.cB = a.B; p.cHbox = Hbox; p.cBbox = Bbox; p.c_tiles_h = cdiv(a.H, Hbox); p.c_cpc = a.Cin / TC_BK;
    p.m_tiles = cdiv(a.B, Bbox) * p.c_tiles_h;
    p.a_bytes = a.W * Hbox * Bbox * TC_BK * 4;
    p.seg_chunks = TC_CONV_SEG_CHUNKS;
    p.resident = tc_resident(w.BN, w.n_tiles, w.k_chunks) ? 1 : 0;
    p.dbg = g_tc_trace;
    p.passes = a.passes == 1 ? 1 : 3;
    const int fixed = 1024 + TC_BAR_BYTES + TC_EPI_WARPS * TC_MAX_BN_RESIDENT * 4 + TC_BM * w.BN * 4 +
                      (p.resident ? w.k_chunks * w.BN * 256 : 0);
    const int sb = TC_A_TILE_BYTES + (p.resident ? 0 : w.BN * (p.passes == 1 ? 128 : 256));   // tf32 mode streams hi only
    p.stages = std::min(TC_MAX_STAGES, (TC_SMEM_LIMIT - fixed) / sb);
    AC_REQUIRE(p.stages >= 2, "conv3x3_tc: tile too large for shared memory (BN=%d)", w.BN);
    p.tmem_cols = 512;
    p.acc_stages = w.BN <= 64 ? TC_MAX_ACC : 2;
    p.a_slots = std::min(4, (512 - p.acc_stages * w.BN) / 64);
    p.merge = p.stages <= p.a_slots ? 1 : 0;
    if (p.merge) p.a_slots = p.stages;
    const size_t smem = (size_t)p.stages * sb + fixed;
    const int grid = std::min(p.m_tiles * p.n_tiles, kNumSMs);
    AC_TIMED("conv3x3_tc", st);
    auto launch = [&](auto kernel) -> int {
        AC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1] = {pdl_attr()};
        cfg.attrs = at; cfg.numAttrs = 1;
        AC_CUDA(cudaLaunchKernelEx(&cfg, kernel, map, map, p));
        return AC_OK;
    };
    int rc;
    if (a.act == ACT_RELU) rc = launch(gemm_tc_kernel<ACT_RELU, false, false, true>);
    else if (a.act == ACT_NONE) rc = launch(gemm_tc_kernel<ACT_NONE, false, false, true>);
    else { set_error("conv3x3_tc: activation %d not supported", a.act); return AC_ERR_ARG; }
    if (rc != AC_OK) return rc;
    AC_LAUNCHED("gemm_tc_kernel<conv3x3>");
    return AC_OK;
}

}  // namespace ac

// Diagnostic / test entry: 3x3 pad-1 convolution + per-channel scale/bias + activation (NHWC fp32).
// w_dev is the PyTorch Conv2d weight [Cout, Cin, 3, 3]; scale/bias [Cout] nullable; act 0 none, 2 relu.
extern "C" int ac_conv3x3(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev,
                          float* out_dev, int B, int H, int W, int Cin, int Cout, int act, void* stream) {
    return ac_conv3x3_p(in_dev, w_dev, scale_dev, bias_dev, out_dev, B, H, W, Cin, Cout, act, 3, stream);
}

extern "C" int ac_conv3x3_p(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev,
                            float* out_dev, int B, int H, int W, int Cin, int Cout, int act, int passes, void* stream) {
    using namespace ac;
    AC_REQUIRE(in_dev && w_dev && out_dev, "ac_conv3x3: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    float *perm = nullptr, *packed = nullptr;
    AC_CUDA(cudaMalloc(&perm, (size_t)Cout * 9 * Cin * sizeof(float)));
    AC_CUDA(cudaMalloc(&packed, tc_packed_floats(Cout, 9 * Cin) * sizeof(float)));
    TcWeight tw;
    int rc = conv3x3_permute_weight(w_dev, perm, Cout, Cin, st);
    if (rc == AC_OK) rc = tc_pack_weight(perm, scale_dev, Cout, 9 * Cin, packed, st, &tw);
    if (rc == AC_OK) {
        Conv3Args a; a.in = in_dev; a.out = out_dev; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
        a.bias = bias_dev; a.tw = &tw; a.act = act; a.passes = passes;
        rc = conv3x3_tc(a, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(perm); cudaFree(packed);
    if (rc == AC_OK && e != cudaSuccess) rc = check_cuda(e, "ac_conv3x3");
    return rc;
}

extern "C" int ac_gemm(const float* A, const float* W, float* C, int M, int N, int K, const float* ascale,
                       int rows_per_group, const float* cscale, const float* cbias, const float* R, int act, int path,
                       void* stream) {
    using namespace ac;
    AC_REQUIRE(A && W && C, "ac_gemm: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    GemmArgs g; g.A = A; g.W = W; g.C = C; g.M = M; g.N = N; g.K = K; g.ascale = ascale;
    g.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; g.cscale = cscale; g.cbias = cbias; g.R = R; g.act = act;
    if (path == 0) return gemm_tn_simt(g, st);
    float* packed = nullptr;
    AC_CUDA(cudaMalloc(&packed, tc_packed_floats(N, K) * sizeof(float)));
    TcWeight tw;
    int rc = tc_pack_weight(W, cscale, N, K, packed, st, &tw);
    if (rc == AC_OK) { g.tw = &tw; rc = gemm_tc(g, st); }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(packed);
    if (rc == AC_OK && e != cudaSuccess) rc = check_cuda(e, "ac_gemm");
    return rc;
}

// Diagnostic: enable (on != 0) / read back the pipeline trace of CTA 0 of the most recent gemm_tc launch.
// out_host receives 8 x 256 clock64 stamps: 0 TMA issue, 1 chunk landed (transform), 2 transform done,
// 3 MMA start, 4 MMA issued, 5 accumulator full (epilogue), 6 epilogue done.
extern "C" int ac_gemm_trace(int on, long long* out_host) {
    using namespace ac;
    if (on && g_tc_trace == nullptr) {
        AC_CUDA(cudaMalloc(&g_tc_trace, 8 * 256 * sizeof(long long)));
        AC_CUDA(cudaMemset(g_tc_trace, 0, 8 * 256 * sizeof(long long)));
    }
    if (out_host != nullptr && g_tc_trace != nullptr) {
        AC_CUDA(cudaDeviceSynchronize());
        AC_CUDA(cudaMemcpy(out_host, g_tc_trace, 8 * 256 * sizeof(long long), cudaMemcpyDeviceToHost));
        AC_CUDA(cudaMemset(g_tc_trace, 0, 8 * 256 * sizeof(long long)));
    }
    if (!on && g_tc_trace != nullptr) { cudaFree(g_tc_trace); g_tc_trace = nullptr; }
    return AC_OK;
}
