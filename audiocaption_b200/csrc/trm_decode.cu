// Transformer caption decoder: KV-cached greedy and beam-search decoding, fp32.
//
// Replaces the reference's `TransformerDecoder.forward` (captioning/models/transformer_decoder.py:80-103,
// HF copy captioning/models/hf_wrapper.py:1045-1068) driven by `CaptionModel.stepwise_forward`
// (captioning/models/base.py:152-218) and `CaptionModel.beam_search` (:254-361) with the
// Transformer glue of captioning/models/transformer_model.py:34-86.
//
// The reference re-evaluates the whole prefix (and re-projects the audio memory) at every step.
// In eval mode that is mathematically identical to:
//   once per call : P = LayerNorm(ReLU(attn_emb W0^T + b0));  K_l, V_l = P Wkv_l^T + bkv_l   (GEMMs)
//   per step      : one token per row through the 2 post-norm layers with cached self-attention
//                   K/V, cross-attention against K_l/V_l, FFN, classifier.
// Clips are independent, so one thread-block CLUSTER owns a small group of rows for the whole decode (all
// `max_len` steps): no grid-wide synchronisation, no host round trip per token, a single launch.  A step is
// bound by streaming the 12.4 MB of weights out of L2.  Two layouts:
//   * head-split (default; greedy_heads_kernel / beam_heads_kernel further down): a cluster of 4 CTAs, CTA h owns attention
//     head h; projections are split over columns where the result can stay local and over K where it must be shared, so a
//     token step needs 7 cluster barriers and every CTA works in every phase; greedy decodes 2 clips per cluster;
//   * column-split (greedy_kernel / beam_kernel, the round-1 layout, kept as fallback): the P CTAs of a cluster each stream
//     1/P of the output columns of every projection / FFN / classifier GEMV and push their slice of the result into every
//     CTA's shared memory; the cheap per-row work (attention over the caches, LayerNorm, log-softmax / arg-max / top-k) is
//     done redundantly by every CTA, so nothing but the GEMV slices is exchanged.  Weights are stored transposed ([K][N])
//     at pack time so that thread n streams column n with coalesced loads while the activations are broadcast from
//     shared memory.
//
// Masks: causal by construction (only positions <= t are cached); `tgt_key_padding_mask`
// = (prefix token == <pad>) and `memory_key_padding_mask` = (frame >= attn_emb_len) are applied
// as -inf before the softmax exactly as nn.MultiheadAttention merges them.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"
#include "decode_common.cuh"
#include "gemm.cuh"

namespace ac {

constexpr int kBeamCluster = 2;     // beam search: CTAs per clip (rows = beams of ONE clip)
constexpr int kGreedyCluster = 2;   // greedy: CTAs per cluster ...
constexpr int kGreedyClips = 1;     // ... which decodes this many clips at once (rows = clips)
// Small batches leave SMs idle at 2 CTAs per clip: spread each clip's 12.4 MB of weights per step over 4 CTAs while
// 4 x batch still fits the 148 SMs (AC_TRM_CLUSTER overrides for measurements).
static int trm_cluster_size(int clips, int dflt) {
    if (const char* e = getenv("AC_TRM_CLUSTER")) { const int p = atoi(e); if (p == 1 || p == 2 || p == 4 || p == 8) return p; }
    return clips <= kNumSMs / 4 ? 4 : dflt;
}
constexpr int D = 256;          // d_model
constexpr int NH = 4;           // heads
constexpr int HD = 64;          // head dim
constexpr int kMaxKeys = 128;   // max(t_mem, max_len)
constexpr int kMaxLen = 64;

struct LayerW {
    const float *sa_in_wt, *sa_in_b, *sa_out_wt, *sa_out_b;   // [D][3D], [3D], [D][D], [D]
    const float *ca_q_wt, *ca_q_b, *ca_out_wt, *ca_out_b;     // [D][D] ...
    const float *ff1_wt, *ff1_b, *ff2_wt, *ff2_b;             // [D][F], [F], [F][D], [D]
    const float *n1_g, *n1_b, *n2_g, *n2_b, *n3_g, *n3_b;
    const float *ca_kv_w, *ca_kv_b;                            // [2D][D] (original layout, for the GEMM), [2D]
};

constexpr int kMaxLayers = 4;

struct DecW {
    const float* emb;     // [V][D]
    const float* pe;      // [pe_len][D]
    const float* cls_wt;  // [D][Vp], Vp = vocab rounded up to 4 (zero-padded columns)
    const float *proj_w, *proj_b, *proj_ln_g, *proj_ln_b;   // attn_proj: [D][attn_emb_dim] original layout
    LayerW layer[kMaxLayers];
    int nlayers, dff, vocab, attn_emb_dim, pe_len;
};

// ------------------------------------------------------------------------------------ helpers
// x[r][:] = LayerNorm(x[r][:] + add[r][:]) * g + b   (eps 1e-5, biased variance), one warp per row
template <int R>
__device__ __forceinline__ void add_layernorm(float* x, const float* add, const float* __restrict__ g,
                                              const float* __restrict__ b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        float v[D / 32];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            v[i] = x[r * D + lane + 32 * i] + add[r * D + lane + 32 * i];
            s += v[i];
        }
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            const int c = lane + 32 * i;
            x[r * D + c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
        }
    }
}

// softmax over n keys for every (row, head); sc layout [R][NH][kMaxKeys]
template <int R>
__device__ __forceinline__ void softmax_rows(float* sc, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < R * NH; i += kWarps) {
        float* p = sc + i * kMaxKeys;
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, p[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) { float e = expf(p[j] - m); p[j] = e; s += e; }
        s = warp_sum(s);
        const float inv = 1.0f / s;
        for (int j = lane; j < n; j += 32) p[j] *= inv;
    }
}

// One warp = one (row, head): scores over `nk` keys (lane j owns key j, j + 32, ...: a 64-long dot product against
// the query broadcast from shared memory), soft-max through warp shuffles, then the probability-weighted sum of the
// value rows with lanes across the 64 head dimensions.  No shared-memory score buffer, no block barrier.
// kptr(j) / vptr(j) give key j's K / V head slice (64 floats, 16-byte aligned); masked(j) keys get -inf.
constexpr int kStageStride = 2 * D + 4;   // padded frame stride of the staged cross K|V: conflict-free LDS.128 per key
template <class KP, class VP, class MK>
__device__ __forceinline__ void attend_head(const float* q, float scale, int nk, KP kptr, VP vptr, MK masked, float* out) {
    const int lane = threadIdx.x & 31;
    constexpr int NC = kMaxKeys / 32;
    float sc[NC];
    const float4* q4 = reinterpret_cast<const float4*>(q);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int j = lane + 32 * i;
        float sv = -INFINITY;
        if (i * 32 < nk && j < nk && !masked(j)) {
            const float4* k4 = reinterpret_cast<const float4*>(kptr(j));
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < HD / 4; ++d) {
                const float4 kv = k4[d], qv = q4[d];
                acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc);
                acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
            }
            sv = acc * scale;
        }
        sc[i] = sv;
    }
    float m = sc[0];
#pragma unroll
    for (int i = 1; i < NC; ++i) m = fmaxf(m, sc[i]);
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) { sc[i] = (i * 32 < nk) ? expf(sc[i] - m) : 0.f; sum += sc[i]; }
    const float inv = 1.0f / warp_sum(sum);
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        if (i * 32 < nk) {
            const float pi = sc[i] * inv;
            const int n = min(32, nk - i * 32);
            for (int jj = 0; jj < n; ++jj) {
                const float pj = __shfl_sync(0xffffffffu, pi, jj);
                const float* v = vptr(i * 32 + jj);
                o0 = fmaf(pj, v[lane], o0);
                o1 = fmaf(pj, v[lane + 32], o1);
            }
        }
    }
    out[lane] = o0;
    out[lane + 32] = o1;
}

struct DecodeArgs {
    DecW w;
    const float* kv_mem;        // [nlayers][clips][t_mem][2D]  (K | V per frame)
    int n_clips;
    const int64_t* mem_len;     // [clips]
    float* kv_cache;            // [clips][nlayers][2][max_len][R][D]
    float* logits_ws;           // [clips][R][V] scratch (beam) ...
    int t_mem, max_len, start_idx, end_idx, pad_idx;
    // greedy outputs
    int64_t* seq;               // [clips][max_len]
    float* logprob;             // nullable [clips][max_len]
    float* logit_out;           // nullable [clips][max_len][V]
    float* embed_out;           // nullable [clips][max_len][D]
    // beam
    int beam; float temp;
    // scheduled-sampling decode (training, base.py:152-170 with mode == "train"): no early stop, no <end> forcing, and
    // where forced[clip][t] >= 0 that token is emitted (and fed back) instead of the arg-max
    const int64_t* forced;      // nullable [clips][max_len]
    int train_mode;
    long long* dbg;             // optional phase trace of CTA 0 (ac_trm_trace): clock64 stamps
    int kv_in_smem;             // the launch reserved shared memory for the cross-attention K/V
};

#define AC_DEC_STAMP(id)                                                                                         \
    do {                                                                                                         \
        if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && t == 5 && (id) < 64) a.dbg[id] = clock64(); \
    } while (0)

// One decoder step for the R rows of this cluster: token ids `words[r]` at position t.  Rows are grouped RPC
// per clip: row r belongs to clip `clip + r / RPC` (greedy: RPC = 1, every row is its own clip; beam: RPC = R).
// anc[r][j] = row (of this cluster, same clip) whose cache slot holds row r's ancestor at position j.
// On return s_x holds the final hidden states [R][D]; row r's logits are written to logits[r] ([V], global).
template <int R, int RPC>
__device__ void decoder_step(const DecodeArgs& a, int clip, int t, const int* words, const int (*anc)[kMaxLen],
                             const unsigned char (*padflag)[8], float* s_x, float* s_q, float* s_att, float* s_h,
                             float* s_sc, float* s_part, float* s_c, float* const* logits, const float* s_kv = nullptr) {
    // Buffers written through distributed shared memory by the peer CTA (GEMV outputs): s_h, s_q, s_c.  Each is
    // only ever the output of a GEMV whose surrounding cluster.sync() interval does not touch it otherwise.
    const DecW& W = a.w;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), P = (int)cluster.num_blocks();
    const int tid = threadIdx.x, warp = tid >> 5;
    auto clip_of = [&](int r) { return min(clip + r / RPC, a.n_clips - 1); };   // rows past the batch replay the last clip
    auto n_mem_of = [&](int r) { return min((int)min((int64_t)a.t_mem, a.mem_len[clip_of(r)]), a.t_mem); };
    AC_DEC_STAMP(0);
    // embedding * sqrt(d) + positional encoding
    for (int i = tid; i < R * D; i += kThreads) {
        const int r = i / D, f = i - r * D;
        s_x[i] = __ldg(W.emb + (size_t)words[r] * D + f) * 16.0f + __ldg(W.pe + (size_t)t * D + f);
    }
    __syncthreads();

    for (int l = 0; l < W.nlayers; ++l) {
        const LayerW& L = W.layer[l];
        // cache of row r: [clip][layer][K|V][position][RPC][D]
        auto kcache = [&](int r, int kv) {
            return a.kv_cache + ((((size_t)clip_of(r) * W.nlayers + l) * 2 + kv) * a.max_len) * RPC * D;
        };
        // ---- self attention
        AC_DEC_STAMP(1 + 12 * l);
        matvec_t<R>(L.sa_in_wt, L.sa_in_b, s_x, D, s_h, 3 * D, 3 * D, D, false, s_part, rank, P);
        AC_DEC_STAMP(2 + 12 * l);
        cluster.sync();
        AC_DEC_STAMP(3 + 12 * l);
        for (int i = tid; i < R * D; i += kThreads) {
            const int r = i / D, f = i - r * D;
            s_c[i] = s_h[r * 3 * D + f] * 0.125f;
            if (clip + r / RPC < a.n_clips) {   // rows past the batch only read the clip they replay
                kcache(r, 0)[((size_t)t * RPC + r % RPC) * D + f] = s_h[r * 3 * D + D + f];
                kcache(r, 1)[((size_t)t * RPC + r % RPC) * D + f] = s_h[r * 3 * D + 2 * D + f];
            }
        }
        __syncthreads();   // also orders the cache writes before the reads below (same CTA)
        const int nk = t + 1;
        for (int task = warp; task < R * NH; task += kWarps) {
            const int r = task / NH, hh = task - r * NH;
            const float* kc = kcache(r, 0) + hh * HD;
            const float* vc = kcache(r, 1) + hh * HD;
            attend_head(s_c + r * D + hh * HD, 1.0f, nk,
                        [&](int j) { return kc + ((size_t)j * RPC + anc[r][j] % RPC) * D; },
                        [&](int j) { return vc + ((size_t)j * RPC + anc[r][j] % RPC) * D; },
                        [&](int j) { return padflag[j][anc[r][j]] != 0; }, s_att + r * D + hh * HD);
        }
        __syncthreads();
        AC_DEC_STAMP(4 + 12 * l);
        matvec_t<R>(L.sa_out_wt, L.sa_out_b, s_att, D, s_q, D, D, D, false, s_part, rank, P);
        cluster.sync();
        AC_DEC_STAMP(5 + 12 * l);
        add_layernorm<R>(s_x, s_q, L.n1_g, L.n1_b);
        __syncthreads();
        // ---- cross attention over the projected audio memory
        AC_DEC_STAMP(6 + 12 * l);
        matvec_t<R>(L.ca_q_wt, L.ca_q_b, s_x, D, s_c, D, D, D, false, s_part, rank, P);
        cluster.sync();
        AC_DEC_STAMP(7 + 12 * l);
        // cross-attention K | V of row r's clip: staged shared-memory copy ([layer][clip of the cluster][frame], padded
        // frame stride) when the kernel made one, else global memory
        for (int task = warp; task < R * NH; task += kWarps) {
            const int r = task / NH, hh = task - r * NH;
            const float* km; int stride;
            if (s_kv != nullptr) { km = s_kv + (((size_t)l * (R / RPC) + r / RPC) * a.t_mem) * kStageStride; stride = kStageStride; }
            else { km = a.kv_mem + (((size_t)l * a.n_clips + clip_of(r)) * a.t_mem) * 2 * D; stride = 2 * D; }
            const int n_mem = n_mem_of(r);
            attend_head(s_c + r * D + hh * HD, 0.125f, a.t_mem,
                        [&](int j) { return km + (size_t)j * stride + hh * HD; },
                        [&](int j) { return km + (size_t)j * stride + D + hh * HD; },
                        [&](int j) { return j >= n_mem; }, s_att + r * D + hh * HD);
        }
        __syncthreads();
        AC_DEC_STAMP(8 + 12 * l);
        matvec_t<R>(L.ca_out_wt, L.ca_out_b, s_att, D, s_q, D, D, D, false, s_part, rank, P);
        cluster.sync();
        AC_DEC_STAMP(9 + 12 * l);
        add_layernorm<R>(s_x, s_q, L.n2_g, L.n2_b);
        __syncthreads();
        // ---- feed forward
        AC_DEC_STAMP(10 + 12 * l);
        matvec_t<R>(L.ff1_wt, L.ff1_b, s_x, D, s_h, W.dff, W.dff, D, true, s_part, rank, P);
        cluster.sync();
        AC_DEC_STAMP(11 + 12 * l);
        matvec_t<R>(L.ff2_wt, L.ff2_b, s_h, W.dff, s_q, D, D, W.dff, false, s_part, rank, P);
        cluster.sync();
        AC_DEC_STAMP(12 + 12 * l);
        add_layernorm<R>(s_x, s_q, L.n3_g, L.n3_b);
        __syncthreads();
    }
    AC_DEC_STAMP(40);
    // ---- classifier (no bias): each thread owns column quads of the zero-padded [D][Vp] weight
    const int V = W.vocab, VC = (V + 3) >> 2;
    const int vc0 = (int)((int64_t)VC * rank / P), vc1 = (int)((int64_t)VC * (rank + 1) / P);
    for (int c = vc0 + tid; c < vc1; c += kThreads) {
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const float4* w = reinterpret_cast<const float4*>(W.cls_wt) + c;
#pragma unroll 8
        for (int k = 0; k < D; ++k) {
            const float4 wv = __ldg(w + (size_t)k * VC);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float xv = s_x[r * D + k];
                acc[r][0] = fmaf(wv.x, xv, acc[r][0]); acc[r][1] = fmaf(wv.y, xv, acc[r][1]);
                acc[r][2] = fmaf(wv.z, xv, acc[r][2]); acc[r][3] = fmaf(wv.w, xv, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * c + q < V && clip + r / RPC < a.n_clips) logits[r][4 * c + q] = acc[r][q];
    }
    cluster.sync();   // both halves of the logits (global memory) are visible to both CTAs
    AC_DEC_STAMP(41);
}

// block-wide (max value, lowest index) reduction
// cross-attention K/V of the cluster's clips staged in shared memory when it fits (128 KB for 2 layers x 32 frames)
constexpr int kDecSmemLimit = 224 * 1024;
static size_t kv_smem_floats(int nlayers, int clips_per_cluster, int t_mem) {
    return (size_t)nlayers * clips_per_cluster * t_mem * kStageStride;
}
template <int NCLIPS>
__device__ __forceinline__ const float* stage_cross_kv(const DecodeArgs& a, int clip0, float* dst) {
    if (!a.kv_in_smem) return nullptr;
    const int per_clip4 = a.t_mem * 2 * D / 4;
    for (int l = 0; l < a.w.nlayers; ++l)
        for (int c = 0; c < NCLIPS; ++c) {
            const int clip = min(clip0 + c, a.n_clips - 1);
            const float4* src = reinterpret_cast<const float4*>(a.kv_mem + (((size_t)l * a.n_clips + clip) * a.t_mem) * 2 * D);
            float* d = dst + (((size_t)l * NCLIPS + c) * a.t_mem) * kStageStride;
            for (int i = threadIdx.x; i < per_clip4; i += kThreads) {
                const int frame = i / (2 * D / 4), q = i - frame * (2 * D / 4);
                reinterpret_cast<float4*>(d + (size_t)frame * kStageStride)[q] = __ldg(src + i);
            }
        }
    __syncthreads();
    return dst;
}

constexpr size_t dec_smem_floats(int R) { return (size_t)R * (D + D + D + 1024 + NH * kMaxKeys + 4096 + D); }

// ------------------------------------------------------------------------------------ greedy
// One cluster decodes G = kGreedyClips clips at once (row r = clip blockIdx.x / P * G + r).
template <int G>
__global__ void __launch_bounds__(kThreads)
greedy_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* s_x = smem; float* s_q = s_x + G * D; float* s_att = s_q + G * D; float* s_h = s_att + G * D;
    float* s_sc = s_h + G * 1024;
    float* s_part = s_sc + G * NH * kMaxKeys;
    float* s_c = s_part + G * 4096;
    __shared__ int s_anc[G][kMaxLen];
    __shared__ unsigned char s_pad[kMaxLen][8];
    __shared__ float s_rv[kWarps];
    __shared__ int s_ri[kWarps];
    __shared__ int s_word[G];
    __shared__ float* s_logits[G];
    cg::cluster_group cluster = cg::this_cluster();
    const int P = (int)cluster.num_blocks();
    const int clip0 = (blockIdx.x / P) * G, tid = threadIdx.x;
    const bool writer = cluster.block_rank() == 0;   // every CTA of the cluster computes the same results
    const int V = a.w.vocab;
    const float* s_kv = stage_cross_kv<G>(a, clip0, s_c + G * D);
    for (int i = tid; i < G * kMaxLen; i += kThreads) s_anc[i / kMaxLen][i % kMaxLen] = i / kMaxLen;
    int word[G]; bool finished[G]; bool valid[G];
#pragma unroll
    for (int r = 0; r < G; ++r) { word[r] = a.start_idx; valid[r] = clip0 + r < a.n_clips; finished[r] = !valid[r]; }
    cluster.sync();   // every CTA of the cluster is resident before the first GEMV writes into its peers' shared memory
    // The reference keeps running finished rows (input forced to <end>) until EVERY row of the batch
    // has finished and records their logits/embeds; when those outputs are requested we do the same
    // for all max_len steps (a superset: past the reference's break its buffers are uninitialised).
    const bool full_outputs = a.logit_out != nullptr || a.embed_out != nullptr;
    for (int t = 0; t < a.max_len; ++t) {
        bool all_done = true;
#pragma unroll
        for (int r = 0; r < G; ++r) all_done = all_done && finished[r];
        if (all_done && !full_outputs) {   // rows that emitted <end> keep <end> (base.py:161-168)
            if (tid < G && writer && clip0 + tid < a.n_clips) a.seq[(size_t)(clip0 + tid) * a.max_len + t] = a.end_idx;
            continue;
        }
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < G; ++r) {
                const int clip = min(clip0 + r, a.n_clips - 1);
                s_pad[t][r] = (word[r] == a.pad_idx); s_word[r] = word[r];
                s_logits[r] = a.logit_out ? a.logit_out + ((size_t)clip * a.max_len + t) * V
                                          : a.logits_ws + (size_t)clip * V;
            }
        }
        __syncthreads();
        decoder_step<G, 1>(a, clip0, t, s_word, s_anc, s_pad, s_x, s_q, s_att, s_h, s_sc, s_part, s_c, s_logits, s_kv);
#pragma unroll
        for (int r = 0; r < G; ++r) {
            if (a.embed_out && writer && valid[r] && tid < D)
                a.embed_out[((size_t)(clip0 + r) * a.max_len + t) * D + tid] = s_x[r * D + tid];
            // log-softmax + argmax (first maximum wins, as torch.max on CPU)
            const float* logits = s_logits[r];
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int n = tid; n < V; n += kThreads) {
                float v = logits[n];
                if (v > best) { best = v; bi = n; }
            }
            block_argmax(best, bi, s_rv, s_ri);
            float se = 0.f;
            for (int n = tid; n < V; n += kThreads) se += expf(logits[n] - best);
            se = block_sum(se, s_rv);
            word[r] = finished[r] ? a.end_idx : bi;
            if (a.forced != nullptr && valid[r]) {
                const int64_t f = a.forced[(size_t)(clip0 + r) * a.max_len + t];
                if (f >= 0) word[r] = (int)f;
            }
            if (tid == 0 && writer && valid[r]) {
                a.seq[(size_t)(clip0 + r) * a.max_len + t] = word[r];
                if (a.logprob) a.logprob[(size_t)(clip0 + r) * a.max_len + t] = -logf(se);
            }
            finished[r] = !valid[r] || (!a.train_mode && (finished[r] || (word[r] == a.end_idx)));
        }
        AC_DEC_STAMP(42);
    }
}

// ------------------------------------------------------------------------------------ greedy, one CTA per head
// Second greedy layout (the default when it fits): a cluster of NH = 4 CTAs decodes G clips, CTA h OWNS attention head h.
//   * projections feeding attention (self q|k|v, cross q) and the first FFN layer are split over output columns so that
//     CTA h computes exactly head h's 64 columns (a quarter of the hidden units): the results stay LOCAL, no exchange;
//   * the projections that follow (attention out, second FFN layer) are split over K -- CTA h multiplies its own 64
//     (256) inputs with the matching weight rows -- and the four partial [G][256] vectors meet in every CTA's shared
//     memory (one distributed-shared-memory push + cluster.sync), where the residual add + LayerNorm is done redundantly;
//   * the classifier is split over the vocabulary; each CTA reduces its slice to (max, arg-max, sum-exp) per row and
//     only those three numbers are exchanged.
// 7 cluster barriers per token step instead of 13, every CTA busy in every phase (attention included), the self- and
// cross-attention K/V of the CTA's head live in shared memory, and with G = 2 the 12.4 MB of weights are streamed once
// per TWO clips: half the L2 traffic of greedy_kernel<1> (which is what bounds it: 794 MB per step at 64 clips).
constexpr int kHeadStride = 2 * HD + 4;       // K | V of one head per key, padded: conflict-free LDS.128 per key
constexpr int kHeadPart = 8192;               // floats of split-K scratch

// Slice GEMV: out[r][j], j in [0, 4 NQ) = sum_{k0 <= k < k1} xin[r][k - k0] * Wt[k][column quad colq(j / 4)].
// thread = (K-slice, column quad): independent 128-bit loads, partial sums through `part`, fixed-order reduction.
template <int G, class ColQ, class Epi>
__device__ __forceinline__ void gemv_slice(const float* __restrict__ Wt, int ldw4, ColQ colq, int NQ, int k0, int k1,
                                           const float* xin, int ldx, float* part, Epi epi) {
    const int tid = threadIdx.x;
    const int Nl = NQ * 4, K = k1 - k0;
    const int KS = max(1, min(min(kThreads / NQ, kHeadPart / (G * Nl)), K));
    const int kslice = ((K + KS - 1) / KS + 3) & ~3;     // multiple of 4: a slice's activations are read as float4 (ldx % 4 == 0)
    const int s = tid / NQ, cl = tid - s * NQ;
    if (s < KS) {
        const int ka = k0 + s * kslice, kb = min(k1, ka + kslice);
        float acc[G][4];
#pragma unroll
        for (int r = 0; r < G; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const float4* w = reinterpret_cast<const float4*>(Wt) + colq(cl);
        constexpr int U = G <= 2 ? 8 : 4;                // independent 128-bit loads in flight per thread (64 registers)
        int k = ka;
        for (; k + U <= kb; k += U) {
            float4 wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) wv[u] = __ldg(w + (size_t)(k + u) * ldw4);
            if constexpr (G <= 2) {                       // activations as float4: a quarter of the LDS instructions
#pragma unroll
                for (int r = 0; r < G; ++r) {
#pragma unroll
                    for (int u4 = 0; u4 < U / 4; ++u4) {
                        const float4 xq = *reinterpret_cast<const float4*>(xin + r * ldx + (k + 4 * u4 - k0));
                        const float xs[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 ww = wv[4 * u4 + e];
                            acc[r][0] = fmaf(ww.x, xs[e], acc[r][0]); acc[r][1] = fmaf(ww.y, xs[e], acc[r][1]);
                            acc[r][2] = fmaf(ww.z, xs[e], acc[r][2]); acc[r][3] = fmaf(ww.w, xs[e], acc[r][3]);
                        }
                    }
                }
            } else {                                      // (more rows: the extra live registers spill)
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int r = 0; r < G; ++r) {
                        const float xv = xin[r * ldx + (k + u - k0)];
                        acc[r][0] = fmaf(wv[u].x, xv, acc[r][0]); acc[r][1] = fmaf(wv[u].y, xv, acc[r][1]);
                        acc[r][2] = fmaf(wv[u].z, xv, acc[r][2]); acc[r][3] = fmaf(wv[u].w, xv, acc[r][3]);
                    }
                }
            }
        }
        if (k < kb) {                                     // tail: the remaining (< U) loads, again all in flight
            float4 wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) wv[u] = k + u < kb ? __ldg(w + (size_t)(k + u) * ldw4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k + u < kb) {
#pragma unroll
                    for (int r = 0; r < G; ++r) {
                        const float xv = xin[r * ldx + (k + u - k0)];
                        acc[r][0] = fmaf(wv[u].x, xv, acc[r][0]); acc[r][1] = fmaf(wv[u].y, xv, acc[r][1]);
                        acc[r][2] = fmaf(wv[u].z, xv, acc[r][2]); acc[r][3] = fmaf(wv[u].w, xv, acc[r][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < G; ++r)
            *reinterpret_cast<float4*>(part + ((size_t)s * G + r) * Nl + 4 * cl) =
                make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    __syncthreads();
    for (int i = tid; i < G * Nl; i += kThreads) {
        const int r = i / Nl, j = i - r * Nl;
        float v = 0.f;
        for (int q = 0; q < KS; ++q) v += part[((size_t)q * G + r) * Nl + j];
        epi(r, j, v);
    }
}

// x[r] = LayerNorm(x[r] + bias + sum_p red[p][r]) * g + b   (the four CTAs' K-split partial products), warp per row
template <int G>
__device__ __forceinline__ void reduce_add_layernorm(float* x, const float* red, const float* __restrict__ bias,
                                                     const float* __restrict__ g, const float* __restrict__ b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < G) {
        const int r = warp;
        float v[D / 32];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            const int c = lane + 32 * i;
            const float add = (red[(0 * G + r) * D + c] + red[(1 * G + r) * D + c]) +
                              (red[(2 * G + r) * D + c] + red[(3 * G + r) * D + c]);
            v[i] = x[r * D + c] + (add + bias[c]);
            s += v[i];
        }
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            const int c = lane + 32 * i;
            x[r * D + c] = (v[i] - mean) * rstd * g[c] + b[c];
        }
    }
}

__host__ __device__ constexpr int head_par_stride(int dff) { return 3 * HD + HD + dff / NH + 9 * D; }
struct HeadSmem { size_t x, qkv, att, hid, part, red, logit, cls, kv, self, par, total; int self_global, self_stride; };
// self_global: the self-attention K/V of the CTA's head live in its private slice of the global cache workspace (stride
// 2 HD, read back through L2) instead of shared memory -- for the row counts whose cache does not fit (beam 4 and 5)
static HeadSmem head_smem(int G, int clips, int nlayers, int dff, int vocab, int t_mem, int max_len, bool self_global = false) {
    HeadSmem h; size_t o = 0;
    auto take = [&](size_t n) { const size_t at = o; o += (n + 3) / 4 * 4; return at; };
    h.x = take((size_t)G * D); h.qkv = take((size_t)G * 3 * HD); h.att = take((size_t)G * HD);
    h.hid = take((size_t)G * (dff / NH)); h.part = take(kHeadPart); h.red = take((size_t)2 * NH * G * D);
    h.logit = take((size_t)G * (cdiv((vocab + 3) / 4, NH) * 4)); h.cls = take((size_t)2 * NH * G * 4);
    h.kv = take((size_t)nlayers * clips * t_mem * kHeadStride);
    h.par = take((size_t)nlayers * head_par_stride(dff));
    h.self = self_global ? 0 : take((size_t)nlayers * G * max_len * kHeadStride);
    h.self_global = self_global ? 1 : 0; h.self_stride = self_global ? 2 * HD : kHeadStride;
    h.total = o;
    return h;
}

struct HeadBufs { float *x, *qkv, *att, *hid, *part, *red, *logit, *kv, *self; int self_stride; float* par; };

// Biases and LayerNorm parameters of every layer, staged once per CTA (the streamed weights evict them from L1, and an L2
// round trip at the end of each GEMV / in front of each LayerNorm is ~10 % of a token step).  Per layer:
// [sa_in_b of my head: q|k|v 192][ca_q_b of my head 64][ff1_b of my hidden units dff/4][9 x 256: sa_out_b n1_g n1_b
//  ca_out_b n2_g n2_b ff2_b n3_g n3_b]
__device__ __forceinline__ void heads_stage_par(const DecW& W, int h, float* par) {
    const int dffl = W.dff / NH, stride = head_par_stride(W.dff);
    for (int l = 0; l < W.nlayers; ++l) {
        const LayerW& L = W.layer[l];
        float* P = par + (size_t)l * stride;
        for (int i = threadIdx.x; i < stride; i += kThreads) {
            float v;
            if (i < 3 * HD) v = L.sa_in_b[(i >> 6) * D + h * HD + (i & 63)];
            else if (i < 4 * HD) v = L.ca_q_b[h * HD + (i - 3 * HD)];
            else if (i < 4 * HD + dffl) v = L.ff1_b[h * dffl + (i - 4 * HD)];
            else {
                const int j = i - 4 * HD - dffl, which = j / D, c = j - which * D;
                const float* src = which == 0 ? L.sa_out_b : which == 1 ? L.n1_g : which == 2 ? L.n1_b : which == 3 ? L.ca_out_b
                                 : which == 4 ? L.n2_g : which == 5 ? L.n2_b : which == 6 ? L.ff2_b : which == 7 ? L.n3_g : L.n3_b;
                v = src[c];
            }
            P[i] = v;
        }
    }
}

// cross-attention K | V of head h for the cluster's clips [clip0, clip0 + NC): [layer][clip slot][frame][K 64 | V 64 | pad]
template <int NC>
__device__ __forceinline__ void heads_stage_kv(const DecodeArgs& a, int clip0, int h, float* s_kv) {
    for (int l = 0; l < a.w.nlayers; ++l)
        for (int c = 0; c < NC; ++c) {
            const int clip = min(clip0 + c, a.n_clips - 1);      // slots past the batch replay the last clip
            const float4* src = reinterpret_cast<const float4*>(a.kv_mem + (((size_t)l * a.n_clips + clip) * a.t_mem) * 2 * D);
            float* dst = s_kv + ((size_t)(l * NC + c) * a.t_mem) * kHeadStride;
            for (int i = threadIdx.x; i < a.t_mem * 32; i += kThreads) {
                const int frame = i >> 5, q = i & 31;              // q < 16: K quad, else V quad
                const float4 v = __ldg(src + (size_t)frame * (2 * D / 4) + (q >> 4) * (D / 4) + h * (HD / 4) + (q & 15));
                reinterpret_cast<float4*>(dst + (size_t)frame * kHeadStride)[q] = v;
            }
        }
}

// One token step of the G rows of a cluster, executed by the CTA of head h.  Row r: token words[r] at position t; rows are
// grouped RPC per clip (greedy: 1, beam search: the beams of a clip); anc(r, j) = row slot whose cache entry at position
// j belongs to row r's history (greedy: r).  On return every CTA holds the final hidden states in s.x and its own slice
// of the logits in s.logit[r][ldl] (vocabulary quads [vc0, vc0 + VQ)); emit(r, j, v) sees every logit of the slice.
template <int G, int RPC, class Anc, class Emit>
__device__ __forceinline__ void heads_step(const DecodeArgs& a, const HeadBufs& s, int h, int t, const int* words,
                                           const unsigned char (*padflag)[8], const int* nmem, Anc anc, int& xc,
                                           int vc0, int VQ, int ldl, Emit emit) {
    cg::cluster_group cluster = cg::this_cluster();
    const DecW& W = a.w;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int dffl = W.dff / NH, VC = (W.vocab + 3) >> 2;
    constexpr int NC = G / RPC;
    // push one row-slice value into the same slot of all four CTAs
    auto push_red = [&](int r, int j, float v) {
        float* slot = s.red + (((size_t)(xc & 1) * NH + h) * G + r) * D + j;
#pragma unroll
        for (int pr = 0; pr < NH; ++pr) *cluster.map_shared_rank(slot, pr) = v;
    };
    AC_DEC_STAMP(0);
    for (int i = tid; i < G * D; i += kThreads) {
        const int r = i / D, f = i - r * D;
        s.x[i] = __ldg(W.emb + (size_t)words[r] * D + f) * 16.0f + __ldg(W.pe + (size_t)t * D + f);
    }
    __syncthreads();
    for (int l = 0; l < W.nlayers; ++l) {
        const LayerW& L = W.layer[l];
        const float* P = s.par + (size_t)l * head_par_stride(W.dff);          // staged biases / LayerNorm parameters
        const float* P9 = P + 4 * HD + dffl;
        // ---- self attention, head h: q | k | v columns of my head (local), cache row t in shared memory
        AC_DEC_STAMP(1 + 12 * l);
        gemv_slice<G>(L.sa_in_wt, 3 * D / 4, [&](int cl) { return (cl >> 4) * (D / 4) + h * (HD / 4) + (cl & 15); }, 48, 0, D,
                      s.x, D, s.part, [&](int r, int j, float v) {
                          v += P[j];
                          if (j < HD) s.qkv[r * 3 * HD + j] = v;
                          else s.self[((size_t)(l * G + r) * a.max_len + t) * s.self_stride + (j - HD)] = v;
                      });
        __syncthreads();
        AC_DEC_STAMP(3 + 12 * l);
        if (warp < G) {
            const int r = warp;
            const float* kc = s.self + ((size_t)l * G * a.max_len) * s.self_stride;
            attend_head(s.qkv + r * 3 * HD, 0.125f, t + 1,
                        [&](int j) { return kc + ((size_t)anc(r, j) * a.max_len + j) * s.self_stride; },
                        [&](int j) { return kc + ((size_t)anc(r, j) * a.max_len + j) * s.self_stride + HD; },
                        [&](int j) { return padflag[j][anc(r, j)] != 0; }, s.att + r * HD);
        }
        __syncthreads();
        AC_DEC_STAMP(4 + 12 * l);
        gemv_slice<G>(L.sa_out_wt, D / 4, [&](int cl) { return cl; }, D / 4, h * HD, (h + 1) * HD, s.att, HD, s.part, push_red);
        cluster.sync();
        AC_DEC_STAMP(5 + 12 * l);
        reduce_add_layernorm<G>(s.x, s.red + (size_t)(xc & 1) * NH * G * D, P9, P9 + D, P9 + 2 * D);
        ++xc;
        __syncthreads();
        // ---- cross attention, head h
        AC_DEC_STAMP(6 + 12 * l);
        gemv_slice<G>(L.ca_q_wt, D / 4, [&](int cl) { return h * (HD / 4) + cl; }, HD / 4, 0, D, s.x, D, s.part,
                      [&](int r, int j, float v) { s.qkv[r * 3 * HD + j] = v + P[3 * HD + j]; });
        __syncthreads();
        AC_DEC_STAMP(7 + 12 * l);
        if (warp < G) {
            const int r = warp;
            const float* km = s.kv + ((size_t)(l * NC + r / RPC) * a.t_mem) * kHeadStride;
            const int n_mem = nmem[r / RPC];
            attend_head(s.qkv + r * 3 * HD, 0.125f, a.t_mem,
                        [&](int j) { return km + (size_t)j * kHeadStride; },
                        [&](int j) { return km + (size_t)j * kHeadStride + HD; },
                        [&](int j) { return j >= n_mem; }, s.att + r * HD);
        }
        __syncthreads();
        AC_DEC_STAMP(8 + 12 * l);
        gemv_slice<G>(L.ca_out_wt, D / 4, [&](int cl) { return cl; }, D / 4, h * HD, (h + 1) * HD, s.att, HD, s.part, push_red);
        cluster.sync();
        AC_DEC_STAMP(9 + 12 * l);
        reduce_add_layernorm<G>(s.x, s.red + (size_t)(xc & 1) * NH * G * D, P9 + 3 * D, P9 + 4 * D, P9 + 5 * D);
        ++xc;
        __syncthreads();
        // ---- feed forward: my quarter of the hidden units, then their share of the output
        AC_DEC_STAMP(10 + 12 * l);
        gemv_slice<G>(L.ff1_wt, W.dff / 4, [&](int cl) { return h * (dffl / 4) + cl; }, dffl / 4, 0, D, s.x, D, s.part,
                      [&](int r, int j, float v) { s.hid[r * dffl + j] = fmaxf(v + P[4 * HD + j], 0.0f); });
        __syncthreads();
        AC_DEC_STAMP(11 + 12 * l);
        gemv_slice<G>(L.ff2_wt, D / 4, [&](int cl) { return cl; }, D / 4, h * dffl, (h + 1) * dffl, s.hid, dffl, s.part, push_red);
        cluster.sync();
        AC_DEC_STAMP(12 + 12 * l);
        reduce_add_layernorm<G>(s.x, s.red + (size_t)(xc & 1) * NH * G * D, P9 + 6 * D, P9 + 7 * D, P9 + 8 * D);
        ++xc;
        __syncthreads();
    }
    AC_DEC_STAMP(40);
    // ---- classifier (no bias) over my slice of the vocabulary
    gemv_slice<G>(W.cls_wt, VC, [&](int cl) { return vc0 + cl; }, VQ, 0, D, s.x, D, s.part,
                  [&](int r, int j, float v) { s.logit[r * ldl + j] = v; emit(r, j, v); });
    __syncthreads();
    AC_DEC_STAMP(41);
}

// (max, sum exp(x - max)) of every row's logits slice: WPR warps scan a row (each with its own maximum), results in
// sm[r][part] / sse[r][part] (+ the first arg-max in sidx when it is wanted).  n_my = valid columns of the slice.
template <int G, int WPR>
__device__ __forceinline__ void heads_slice_max_sumexp(const float* s_logit, int ldl, int n_my, int col0, float (*sm)[WPR],
                                                       float (*sse)[WPR], int (*sidx)[WPR]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < G * WPR) {
        const int r = warp / WPR, pt = warp - r * WPR;
        const int chunk = (n_my + WPR - 1) / WPR;
        const int n0 = pt * chunk, n1 = min(n_my, n0 + chunk);
        const float* lg = s_logit + r * ldl;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int n = n0 + lane; n < n1; n += 32) {
            const float v = lg[n];
            if (v > best) { best = v; bi = col0 + n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        float se = 0.f;
        for (int n = n0 + lane; n < n1; n += 32) se += expf(lg[n] - best);
        se = warp_sum(se);
        if (lane == 0) { sm[r][pt] = best; sidx[r][pt] = bi; sse[r][pt] = se; }
    }
    __syncthreads();
}

template <int G>
__global__ void __launch_bounds__(kThreads, 1)
greedy_heads_kernel(DecodeArgs a, HeadSmem lay) {
    extern __shared__ __align__(16) float smem[];
    const HeadBufs s{smem + lay.x, smem + lay.qkv, smem + lay.att, smem + lay.hid, smem + lay.part,
                     smem + lay.red, smem + lay.logit, smem + lay.kv, smem + lay.self, lay.self_stride, smem + lay.par};
    float* s_cls = smem + lay.cls;
    __shared__ unsigned char s_pad[kMaxLen][8];
    __shared__ int s_word[G];
    __shared__ int s_nmem[G];
    constexpr int WPR = 32 / G >= 8 ? 8 : 32 / G;            // warps that scan one row's logits slice
    __shared__ float s_cmax[G][WPR], s_cse[G][WPR];
    __shared__ int s_cidx[G][WPR];
    cg::cluster_group cluster = cg::this_cluster();
    const DecW& W = a.w;
    const int h = (int)cluster.block_rank();                 // my head
    const int clip0 = (blockIdx.x / NH) * G, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool writer = h == 0;
    const int V = W.vocab, VC = (V + 3) >> 2;
    const int vc0 = (int)((int64_t)VC * h / NH), vc1 = (int)((int64_t)VC * (h + 1) / NH);   // my vocabulary quads
    const int VQ = vc1 - vc0, ldl = ((VC + NH - 1) / NH) * 4;
    heads_stage_kv<G>(a, clip0, h, s.kv);
    heads_stage_par(a.w, h, s.par);
    if (tid < G) s_nmem[tid] = min((int)min((int64_t)a.t_mem, a.mem_len[min(clip0 + tid, a.n_clips - 1)]), a.t_mem);
    int word[G]; bool finished[G]; bool valid[G];
#pragma unroll
    for (int r = 0; r < G; ++r) { word[r] = a.start_idx; valid[r] = clip0 + r < a.n_clips; finished[r] = !valid[r]; }
    int xc = 0;                                          // exchange counter: s.red / s_cls are ping-pong buffers
    cluster.sync();   // every CTA of the cluster is resident before the first push into its peers' shared memory
    const bool full_outputs = a.logit_out != nullptr || a.embed_out != nullptr;
    for (int t = 0; t < a.max_len; ++t) {
        bool all_done = true;
#pragma unroll
        for (int r = 0; r < G; ++r) all_done = all_done && finished[r];
        if (all_done && !full_outputs) {   // rows that emitted <end> keep <end> (base.py:161-168)
            if (tid < G && writer && clip0 + tid < a.n_clips) a.seq[(size_t)(clip0 + tid) * a.max_len + t] = a.end_idx;
            continue;
        }
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < G; ++r) { s_pad[t][r] = (word[r] == a.pad_idx); s_word[r] = word[r]; }
        }
        __syncthreads();
        heads_step<G, 1>(a, s, h, t, s_word, s_pad, s_nmem, [](int r, int) { return r; }, xc, vc0, VQ, ldl,
                         [&](int r, int j, float v) {
                             const int col = 4 * vc0 + j;
                             if (a.logit_out != nullptr && col < V && valid[r])
                                 a.logit_out[((size_t)(clip0 + r) * a.max_len + t) * V + col] = v;
                         });
        // my slice -> (max, first arg-max, sum exp(x - max)) per row; one warp merges the WPR parts and pushes the three
        // numbers to all four CTAs
        heads_slice_max_sumexp<G, WPR>(s.logit, ldl, min(4 * VQ, V - 4 * vc0), 4 * vc0, s_cmax, s_cse, s_cidx);
        if (warp < G && lane < NH) {
            const int r = warp;
            float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
            for (int q = 0; q < WPR; ++q)          // chunks are in vocabulary order: the first maximum wins
                if (s_cmax[r][q] > best) { best = s_cmax[r][q]; bi = s_cidx[r][q]; }
            float se = 0.f;
#pragma unroll
            for (int q = 0; q < WPR; ++q) se += s_cse[r][q] > 0.f ? s_cse[r][q] * expf(s_cmax[r][q] - best) : 0.f;
            float* slot = cluster.map_shared_rank(s_cls + (((size_t)(xc & 1) * NH + h) * G + r) * 4, lane);
            slot[0] = best; slot[1] = __int_as_float(bi); slot[2] = se;
        }
        cluster.sync();
        {
            const float* c = s_cls + (size_t)(xc & 1) * NH * G * 4;
            ++xc;
#pragma unroll
            for (int r = 0; r < G; ++r) {
                float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
                for (int pr = 0; pr < NH; ++pr) {     // slices are in vocabulary order: the first maximum wins (torch.max on CPU)
                    const float v = c[(pr * G + r) * 4];
                    const int vi = __float_as_int(c[(pr * G + r) * 4 + 1]);
                    if (v > best) { best = v; bi = vi; }
                }
                float se = 0.f;
#pragma unroll
                for (int pr = 0; pr < NH; ++pr) se += c[(pr * G + r) * 4 + 2] * expf(c[(pr * G + r) * 4] - best);
                if (a.embed_out && writer && valid[r] && tid < D)
                    a.embed_out[((size_t)(clip0 + r) * a.max_len + t) * D + tid] = s.x[r * D + tid];
                word[r] = finished[r] ? a.end_idx : bi;
                if (a.forced != nullptr && valid[r]) {
                    const int64_t f = a.forced[(size_t)(clip0 + r) * a.max_len + t];
                    if (f >= 0) word[r] = (int)f;
                }
                if (tid == 0 && writer && valid[r]) {
                    a.seq[(size_t)(clip0 + r) * a.max_len + t] = word[r];
                    if (a.logprob) a.logprob[(size_t)(clip0 + r) * a.max_len + t] = -logf(se);
                }
                finished[r] = !valid[r] || (!a.train_mode && (finished[r] || (word[r] == a.end_idx)));
            }
        }
        AC_DEC_STAMP(42);
    }
    cluster.sync();   // nobody exits while a peer may still push into its shared memory
}

// ------------------------------------------------------------------------------------ beam search, one CTA per head
// The same 4-CTA layout for beam search: a cluster decodes CL clips with R beams each (G = CL * R rows share every weight
// load).  The scoring of base.py:282-304 is distributed over the vocabulary slices: two (max, sum-exp) exchanges give the
// two log-softmax normalisers of every row, each CTA keeps its R best (score, flat index) candidates per clip and the four
// candidate lists are merged identically in every CTA -- the global top-R is a subset of the four local top-R.
template <int R, int CL>
__global__ void __launch_bounds__(kThreads, 1)
beam_heads_kernel(DecodeArgs a, HeadSmem lay) {
    constexpr int G = R * CL;
    static_assert(G <= 8, "pad flags hold 8 rows");
    extern __shared__ __align__(16) float smem[];
    // self-attention cache of my head: shared memory, or (lay.self_global) my private slice of the global cache workspace
    float* self_cache = lay.self_global
        ? a.kv_cache + (size_t)blockIdx.x * ((size_t)a.w.nlayers * G * a.max_len * 2 * HD)
        : smem + lay.self;
    const HeadBufs s{smem + lay.x, smem + lay.qkv, smem + lay.att, smem + lay.hid, smem + lay.part,
                     smem + lay.red, smem + lay.logit, smem + lay.kv, self_cache, lay.self_stride, smem + lay.par};
    float* s_cls = smem + lay.cls;                        // [2][NH][G][4] exchange slots
    __shared__ unsigned char s_pad[kMaxLen][8];
    __shared__ int s_anc[2][G][kMaxLen];                  // row slot (of the cluster) holding the ancestor at a position
    __shared__ int s_seq[2][G][kMaxLen];
    __shared__ int s_words[G];
    __shared__ float s_score[G];
    __shared__ float s_newscore[G];
    __shared__ int s_newidx[G];                           // flat index (beam of the clip) * V + word
    __shared__ float s_lval[G];
    __shared__ int s_lidx[G];
    __shared__ float s_lse1[G], s_lse2[G], s_max[G];
    __shared__ int s_best_seq[CL][kMaxLen];
    __shared__ int s_best_len[CL], s_ndone[CL], s_stop[CL];
    __shared__ float s_best_score[CL];
    __shared__ int s_nmem[CL];
    __shared__ float s_rv[kWarps];
    __shared__ int s_ri[kWarps];
    constexpr int WPR = 32 / G >= 8 ? 8 : 32 / G;
    __shared__ float s_cmax[G][WPR], s_cse[G][WPR];
    __shared__ int s_cidx[G][WPR];
    cg::cluster_group cluster = cg::this_cluster();
    const DecW& W = a.w;
    const int h = (int)cluster.block_rank();
    const int clip0 = (blockIdx.x / NH) * CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = W.vocab, VC = (V + 3) >> 2;
    const int vc0 = (int)((int64_t)VC * h / NH), vc1 = (int)((int64_t)VC * (h + 1) / NH);
    const int VQ = vc1 - vc0, ldl = ((VC + NH - 1) / NH) * 4;
    const int n_my = min(4 * VQ, V - 4 * vc0), col0 = 4 * vc0;
    const float inv_t = 1.0f / a.temp;
    heads_stage_kv<CL>(a, clip0, h, s.kv);
    heads_stage_par(a.w, h, s.par);
    if (tid < CL) {
        s_nmem[tid] = min((int)min((int64_t)a.t_mem, a.mem_len[min(clip0 + tid, a.n_clips - 1)]), a.t_mem);
        s_ndone[tid] = 0; s_best_len[tid] = 0; s_best_score[tid] = -INFINITY;
        s_stop[tid] = clip0 + tid < a.n_clips ? 0 : 1;      // slots past the batch never report
    }
    if (tid < G) { s_words[tid] = a.start_idx; s_score[tid] = 0.f; }
    for (int i = tid; i < G * kMaxLen; i += kThreads) s_anc[0][i / kMaxLen][i % kMaxLen] = i / kMaxLen;
    int xc = 0;
    cluster.sync();   // peers resident before the first distributed-shared-memory push
    int cur = 0;
    // one (value, value) pair per row from every CTA: push, barrier, then every CTA reads the four contributions
    auto exchange2 = [&](float v0, float v1, int r) {      // called by lanes < NH of warp r
        float* slot = cluster.map_shared_rank(s_cls + (((size_t)(xc & 1) * NH + h) * G + r) * 4, lane);
        slot[0] = v0; slot[1] = v1;
    };
    for (int t = 0; t < a.max_len; ++t) {
        if (tid < G) { s_pad[t][tid] = (s_words[tid] == a.pad_idx); s_anc[cur][tid][t] = tid; }
        __syncthreads();
        heads_step<G, R>(a, s, h, t, s_words, s_pad, s_nmem, [&](int r, int j) { return s_anc[cur][r][j]; }, xc, vc0, VQ, ldl,
                         [](int, int, float) {});
        // ---- lp = log_softmax(log_softmax(logit) / temp) + running score   (base.py:282-290), over vocabulary slices
        // first normaliser: lse1 = max + log sum exp(x - max)
        heads_slice_max_sumexp<G, WPR>(s.logit, ldl, n_my, col0, s_cmax, s_cse, s_cidx);
        if (warp < G && lane < NH) {
            const int r = warp;
            float best = -INFINITY;
#pragma unroll
            for (int q = 0; q < WPR; ++q) best = fmaxf(best, s_cmax[r][q]);
            float se = 0.f;
#pragma unroll
            for (int q = 0; q < WPR; ++q) se += s_cse[r][q] > 0.f ? s_cse[r][q] * expf(s_cmax[r][q] - best) : 0.f;
            exchange2(best, se, r);
        }
        cluster.sync();
        if (tid < G) {
            const float* c = s_cls + (size_t)(xc & 1) * NH * G * 4;
            float m = -INFINITY;
#pragma unroll
            for (int pr = 0; pr < NH; ++pr) m = fmaxf(m, c[(pr * G + tid) * 4]);
            float se = 0.f;
#pragma unroll
            for (int pr = 0; pr < NH; ++pr) se += c[(pr * G + tid) * 4 + 1] * expf(c[(pr * G + tid) * 4] - m);
            s_max[tid] = m; s_lse1[tid] = m + logf(se);
        }
        ++xc;
        __syncthreads();
        // second normaliser over y = (x - lse1) / temp, whose maximum is (max - lse1) / temp
        if (warp < G * WPR) {
            const int r = warp / WPR, pt = warp - r * WPR;
            const int chunk = (n_my + WPR - 1) / WPR;
            const int n0 = pt * chunk, n1 = min(n_my, n0 + chunk);
            const float* lg = s.logit + r * ldl;
            const float l1 = s_lse1[r], m2 = (s_max[r] - l1) * inv_t;
            float se = 0.f;
            for (int n = n0 + lane; n < n1; n += 32) se += expf((lg[n] - l1) * inv_t - m2);
            se = warp_sum(se);
            if (lane == 0) s_cse[r][pt] = se;
        }
        __syncthreads();
        if (warp < G && lane < NH) {
            const int r = warp;
            float se = 0.f;
#pragma unroll
            for (int q = 0; q < WPR; ++q) se += s_cse[r][q];
            exchange2(se, 0.f, r);
        }
        cluster.sync();
        if (tid < G) {
            const float* c = s_cls + (size_t)(xc & 1) * NH * G * 4;
            float se = 0.f;
#pragma unroll
            for (int pr = 0; pr < NH; ++pr) se += c[(pr * G + tid) * 4];
            s_lse2[tid] = (s_max[tid] - s_lse1[tid]) * inv_t + logf(se);
        }
        ++xc;
        __syncthreads();
        // my slice's R best candidates of every clip (step 0: the clip's first row only), R rounds of block arg-max;
        // ties go to the lower flat index, as in the single-slice kernel
        for (int c = 0; c < CL; ++c) {
            const int rows = t == 0 ? 1 : R;
            for (int k = 0; k < R; ++k) {
                float best = -INFINITY; int bi = 0x7fffffff;
                for (int i = tid; i < rows * n_my; i += kThreads) {
                    const int rl = i / n_my, n = i - rl * n_my;
                    const int r = c * R + rl;
                    const int flat = rl * V + col0 + n;
                    bool taken = false;
                    for (int q = 0; q < k; ++q) taken |= (s_lidx[c * R + q] == flat);
                    const float v = s_score[r] + ((s.logit[r * ldl + n] - s_lse1[r]) * inv_t - s_lse2[r]);
                    if (!taken && (v > best || (v == best && flat < bi))) { best = v; bi = flat; }
                }
                block_argmax(best, bi, s_rv, s_ri);
                if (tid == 0) { s_lval[c * R + k] = best; s_lidx[c * R + k] = bi; }
                __syncthreads();
            }
        }
        if (warp < G && lane < NH) exchange2(s_lval[warp], __int_as_float(s_lidx[warp]), warp);
        cluster.sync();
        // merge the 4 R candidates of a clip (identical in every CTA) and update the clip's beams
        if (tid < CL && !s_stop[tid]) {
            const int c = tid;
            const float* cs = s_cls + (size_t)(xc & 1) * NH * G * 4;
            bool used[NH * R];
            for (int i = 0; i < NH * R; ++i) used[i] = false;
            for (int k = 0; k < R; ++k) {
                float best = -INFINITY; int bi = 0x7fffffff, bj = -1;
                for (int i = 0; i < NH * R; ++i) {
                    const int pr = i / R, q = i - pr * R;
                    const float v = cs[(pr * G + c * R + q) * 4];
                    const int vi = __float_as_int(cs[(pr * G + c * R + q) * 4 + 1]);
                    if (!used[i] && vi != 0x7fffffff && (v > best || (v == best && vi < bi))) { best = v; bi = vi; bj = i; }
                }
                if (bj >= 0) used[bj] = true;
                s_newscore[c * R + k] = best; s_newidx[c * R + k] = bi;
            }
        }
        ++xc;
        __syncthreads();
        // reorder beams: histories follow their parent (base.py:291-304)
        const int nxt = cur ^ 1;
        for (int i = tid; i < G * kMaxLen; i += kThreads) {
            const int r = i / kMaxLen, j = i % kMaxLen, c = r / R;
            if (s_stop[c]) continue;
            const int parent = c * R + s_newidx[r] / V;
            if (j <= t) s_anc[nxt][r][j] = s_anc[cur][parent][j];
            if (j < t) s_seq[nxt][r][j] = s_seq[cur][parent][j];
            if (j == t) s_seq[nxt][r][j] = s_newidx[r] % V;
        }
        __syncthreads();
        if (tid < CL && !s_stop[tid]) {
            const int c = tid;
            for (int rl = 0; rl < R; ++rl) {
                const int r = c * R + rl;
                const int wd = s_newidx[r] % V;
                float sc = s_newscore[r];
                const bool is_end = (wd == a.end_idx) || (t == a.max_len - 1);
                if (is_end) {
                    const float fs = sc / (float)(t + 1);
                    s_ndone[c]++;
                    if (fs > s_best_score[c]) {   // stable best-first: earlier beam wins ties
                        s_best_score[c] = fs; s_best_len[c] = t + 1;
                        for (int j = 0; j <= t; ++j) s_best_seq[c][j] = s_seq[nxt][r][j];
                    }
                    sc -= 1000.0f;
                }
                s_score[r] = sc; s_words[r] = wd;
            }
            if (s_ndone[c] == R) s_stop[c] = 1;   // equality, as the reference
        }
        __syncthreads();
        // a stopped clip keeps its last histories in BOTH buffers (its rows are still stepped, their results unused)
        for (int i = tid; i < G * kMaxLen; i += kThreads) {
            const int r = i / kMaxLen, j = i % kMaxLen;
            if (s_stop[r / R] && j <= t) { s_anc[nxt][r][j] = s_anc[cur][r][j]; }
        }
        cur = nxt;
        bool all_stop = true;
#pragma unroll
        for (int c = 0; c < CL; ++c) all_stop = all_stop && s_stop[c] != 0;
        __syncthreads();
        if (all_stop) break;
    }
    if (h == 0)
        for (int i = tid; i < CL * a.max_len; i += kThreads) {
            const int c = i / a.max_len, j = i - c * a.max_len;
            if (clip0 + c < a.n_clips) a.seq[(size_t)(clip0 + c) * a.max_len + j] = j < s_best_len[c] ? s_best_seq[c][j] : a.end_idx;
        }
    cluster.sync();   // nobody exits while a peer may still push into its shared memory
}

// ------------------------------------------------------------------------------------ beam search
template <int R>
__global__ void __launch_bounds__(kThreads)
beam_kernel(DecodeArgs a) {
    __shared__ float* s_logits[R];
    extern __shared__ __align__(16) float smem[];
    float* s_x = smem; float* s_q = s_x + R * D; float* s_att = s_q + R * D; float* s_h = s_att + R * D;
    float* s_sc = s_h + R * 1024;
    float* s_part = s_sc + R * NH * kMaxKeys;
    float* s_c = s_part + R * 4096;
    __shared__ int s_anc[2][R][kMaxLen];
    __shared__ unsigned char s_pad[kMaxLen][8];
    __shared__ int s_seq[2][R][kMaxLen];
    __shared__ int s_words[R];
    __shared__ float s_score[R];
    __shared__ float s_newscore[R];
    __shared__ int s_newidx[R];
    __shared__ float s_rv[kWarps];
    __shared__ int s_ri[kWarps];
    __shared__ int s_best_seq[kMaxLen];
    __shared__ int s_best_len, s_ndone, s_stop;
    __shared__ float s_best_score;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int P = (int)cluster.num_blocks();
    const int clip = blockIdx.x / P, tid = threadIdx.x;
    const int V = a.w.vocab;
    float* lp = a.logits_ws + (size_t)clip * R * V;
    const float* s_kv = stage_cross_kv<1>(a, clip, s_c + R * D);
    if (tid < R) { s_words[tid] = a.start_idx; s_score[tid] = 0.f; s_logits[tid] = lp + (size_t)tid * V; }
    if (tid == 0) { s_ndone = 0; s_stop = 0; s_best_len = 0; s_best_score = -INFINITY; }
    for (int i = tid; i < R * kMaxLen; i += kThreads) { s_anc[0][i / kMaxLen][i % kMaxLen] = i / kMaxLen; }
    cluster.sync();   // peers resident before the first distributed-shared-memory write (decode_common.cuh matvec_t)
    int cur = 0;
    for (int t = 0; t < a.max_len; ++t) {
        if (tid < R) s_pad[t][tid] = (s_words[tid] == a.pad_idx);
        if (tid < R) s_anc[cur][tid][t] = tid;
        __syncthreads();
        decoder_step<R, R>(a, clip, t, s_words, s_anc[cur], s_pad, s_x, s_q, s_att, s_h, s_sc, s_part, s_c, s_logits, s_kv);
        // lp = log_softmax(log_softmax(logit) / temp) + running score   (base.py:282-290)
        for (int r = 0; r < R; ++r) {
            float* row = lp + (size_t)r * V;
            float m = -INFINITY; int mi = 0;
            for (int n = tid; n < V; n += kThreads) m = fmaxf(m, row[n]);
            block_argmax(m, mi, s_rv, s_ri);
            float se = 0.f;
            for (int n = tid; n < V; n += kThreads) se += expf(row[n] - m);
            const float lse1 = m + logf(block_sum(se, s_rv));
            const float inv_t = 1.0f / a.temp;
            // second log-softmax over y = (row - lse1) / temp ; its max is (m - lse1) / temp
            const float m2 = (m - lse1) * inv_t;
            float se2 = 0.f;
            for (int n = tid; n < V; n += kThreads) se2 += expf((row[n] - lse1) * inv_t - m2);
            const float lse2 = m2 + logf(block_sum(se2, s_rv));
            const float sc = s_score[r];
            // in-place rewrite of the shared (global) row: each CTA rewrites its own half exactly once
            const int n0 = (int)((int64_t)V * rank / P), n1 = (int)((int64_t)V * (rank + 1) / P);
            cluster.sync();   // every read of the raw logits above is done in both CTAs
            for (int n = n0 + tid; n < n1; n += kThreads) row[n] = sc + ((row[n] - lse1) * inv_t - lse2);
        }
        cluster.sync();
        // top-R of the flattened [rows * V] scores (step 0: row 0 only), R rounds of block arg-max
        const int ncand = (t == 0 ? 1 : R) * V;
        for (int k = 0; k < R; ++k) {
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int n = tid; n < ncand; n += kThreads) {
                bool taken = false;
                for (int q = 0; q < k; ++q) taken |= (s_newidx[q] == n);
                float v = lp[n];
                if (!taken && (v > best || (v == best && n < bi))) { best = v; bi = n; }
            }
            block_argmax(best, bi, s_rv, s_ri);
            if (tid == 0) { s_newscore[k] = best; s_newidx[k] = bi; }
            __syncthreads();
        }
        // reorder beams: histories follow their parent (base.py:291-304)
        const int nxt = cur ^ 1;
        for (int i = tid; i < R * kMaxLen; i += kThreads) {
            const int r = i / kMaxLen, j = i % kMaxLen;
            const int parent = s_newidx[r] / V;
            if (j <= t) s_anc[nxt][r][j] = s_anc[cur][parent][j];
            if (j < t) s_seq[nxt][r][j] = s_seq[cur][parent][j];
            if (j == t) s_seq[nxt][r][j] = s_newidx[r] % V;
        }
        __syncthreads();
        if (tid == 0) {
            for (int r = 0; r < R; ++r) {
                const int wd = s_newidx[r] % V;
                float sc = s_newscore[r];
                const bool is_end = (wd == a.end_idx) || (t == a.max_len - 1);
                if (is_end) {
                    const float fs = sc / (float)(t + 1);
                    s_ndone++;
                    if (fs > s_best_score) {   // stable best-first: earlier beam wins ties
                        s_best_score = fs; s_best_len = t + 1;
                        for (int j = 0; j <= t; ++j) s_best_seq[j] = s_seq[nxt][r][j];
                    }
                    sc -= 1000.0f;
                }
                s_score[r] = sc; s_words[r] = wd;
            }
            if (s_ndone == R) s_stop = 1;   // equality, as the reference
        }
        __syncthreads();
        cur = nxt;
        if (s_stop) break;
    }
    if (rank == 0)
        for (int j = tid; j < a.max_len; j += kThreads)
            a.seq[(size_t)clip * a.max_len + j] = j < s_best_len ? s_best_seq[j] : a.end_idx;
}

// rows [n, D] in place: LayerNorm(x) * g + b, eps 1e-5; one warp per row
__global__ void layernorm_rows_kernel(float* __restrict__ x, const float* __restrict__ g,
                                      const float* __restrict__ b, int64_t n) {
    const int lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= n) return;
    float* p = x + row * D;
    float v[D / 32]; float s = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { v[i] = p[lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < D / 32; ++i) {
        const int c = lane + 32 * i;
        p[c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
}

// in [rows][cols] -> out [cols][ld] (ld >= rows)
__global__ void transpose_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)rows * cols) {
        int r = (int)(i / cols), c = (int)(i % cols);
        out[(size_t)c * ld + r] = in[i];
    }
}

__global__ void transpose2_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)rows * cols) {
        int r = (int)(i / cols), c = (int)(i % cols);
        out[(size_t)c * rows + r] = in[i];
    }
}

}  // namespace ac

struct ac_trm {
    float* blob = nullptr;
    float* tcblob = nullptr;               // tensor-core images of the memory-side weights (gemm.cuh)
    ac::TcWeight proj_tw, kv_tw[ac::kMaxLayers];
    ac::DecW w;
    std::vector<int64_t> want;             // element count of every source tensor, in tensor order
};

namespace ac {
struct TrmWs { size_t proj, kvmem, cache, logits; };
static TrmWs trm_ws(const ac_trm* d, int clips, int rows_per_clip, int t_mem, int max_len) {
    TrmWs s;
    s.proj = align_up((size_t)clips * t_mem * D, 64);
    s.kvmem = align_up((size_t)clips * d->w.nlayers * t_mem * 2 * D, 64);
    s.cache = align_up((size_t)clips * d->w.nlayers * 2 * max_len * rows_per_clip * D, 64);
    s.logits = align_up((size_t)clips * rows_per_clip * d->w.vocab, 64);
    return s;
}

// attn_proj + per-layer cross-attention K/V of the audio memory (once per call)
static int prepare_memory(const ac_trm* d, const float* attn_emb, int clips, int t_mem, float* proj, float* kvmem,
                          cudaStream_t st) {
    const DecW& W = d->w;
    GemmArgs g; g.A = attn_emb; g.W = W.proj_w; g.C = proj; g.M = clips * t_mem; g.N = D; g.K = W.attn_emb_dim;
    g.cbias = W.proj_b; g.act = ACT_RELU; g.tw = d->proj_tw.packed ? &d->proj_tw : nullptr;
    int rc = gemm_tn(g, st); if (rc) return rc;
    int64_t rows = (int64_t)clips * t_mem;
    {
        AC_TIMED("layernorm_rows", st);
        layernorm_rows_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, st>>>(proj, W.proj_ln_g, W.proj_ln_b, rows);
        AC_LAUNCHED("layernorm_rows_kernel");
    }
    for (int l = 0; l < W.nlayers; ++l) {
        // one [clips*t_mem, D] x [D, 2D] GEMM per layer; kvmem layout [layer][clip][t][K | V]
        GemmArgs k; k.A = proj; k.W = W.layer[l].ca_kv_w; k.M = clips * t_mem; k.N = 2 * D; k.K = D;
        k.cbias = W.layer[l].ca_kv_b; k.act = ACT_NONE; k.tw = d->kv_tw[l].packed ? &d->kv_tw[l] : nullptr;
        k.C = kvmem + (size_t)l * clips * t_mem * 2 * D;   // [layer][clip][t][2D]
        rc = gemm_tn(k, st); if (rc) return rc;
    }
    return AC_OK;
}
}  // namespace ac

// (Re)builds every packed / transposed weight image of `d` from the source tensors `t` (asynchronous on `st`, no allocation).
static int trm_fill(ac_trm_t* d, const float* const* t, cudaStream_t st) {
    using namespace ac;
    const std::vector<int64_t>& want = d->want;
    const int nlayers = d->w.nlayers, F = d->w.dff, V = d->w.vocab, attn_emb_dim = d->w.attn_emb_dim;
    size_t off = 0;
    int ti = 0;
    int rc = AC_OK;
    auto plain = [&]() -> const float* {
        float* p = d->blob + off;
        if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(p, t[ti], want[ti] * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy");
        off += align_up((size_t)want[ti], 64); ++ti;
        return p;
    };
    auto transposed = [&](int rows, int cols, const float* src) -> const float* {   // [rows][cols] -> [cols][rows]
        float* p = d->blob + off;
        transpose2_kernel<<<(unsigned)cdiv64((int64_t)rows * cols, 256), 256, 0, st>>>(src, p, rows, cols);
        g_launches++;
        return p;
    };
    DecW& W = d->w;
    W.emb = plain();
    W.pe = plain();
    for (int l = 0; l < nlayers; ++l) {
        LayerW& L = W.layer[l];
        L.sa_in_wt = transposed(3 * D, D, t[ti]); off += align_up((size_t)want[ti], 64); ++ti;
        L.sa_in_b = plain();
        L.sa_out_wt = transposed(D, D, t[ti]); off += align_up((size_t)want[ti], 64); ++ti;
        L.sa_out_b = plain();
        // multihead_attn.in_proj: rows [0,D) = query projection (transposed for the decode kernel),
        // rows [D,3D) = key/value projections (kept [2D][D] for the memory GEMM)
        {
            float* p = d->blob + off;
            transpose2_kernel<<<(unsigned)cdiv64((int64_t)D * D, 256), 256, 0, st>>>(t[ti], p, D, D);
            g_launches++;
            if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(p + D * D, t[ti] + (size_t)D * D, (size_t)2 * D * D * sizeof(float),
                                                             cudaMemcpyDeviceToDevice, st), "copy kv");
            L.ca_q_wt = p; L.ca_kv_w = p + D * D;
            off += align_up((size_t)want[ti], 64); ++ti;
        }
        { const float* b = plain(); L.ca_q_b = b; L.ca_kv_b = b + D; }
        L.ca_out_wt = transposed(D, D, t[ti]); off += align_up((size_t)want[ti], 64); ++ti;
        L.ca_out_b = plain();
        L.ff1_wt = transposed(F, D, t[ti]); off += align_up((size_t)want[ti], 64); ++ti;
        L.ff1_b = plain();
        L.ff2_wt = transposed(D, F, t[ti]); off += align_up((size_t)want[ti], 64); ++ti;
        L.ff2_b = plain();
        L.n1_g = plain(); L.n1_b = plain(); L.n2_g = plain(); L.n2_b = plain(); L.n3_g = plain(); L.n3_b = plain();
    }
    {   // classifier [V][D] -> zero-padded [D][Vp]
        const int Vp = (V + 3) / 4 * 4;
        float* p = d->blob + off;
        if (rc == AC_OK) rc = check_cuda(cudaMemsetAsync(p, 0, (size_t)D * Vp * sizeof(float), st), "memset cls");
        transpose_pad_kernel<<<(unsigned)cdiv64((int64_t)V * D, 256), 256, 0, st>>>(t[ti], p, V, D, Vp);
        g_launches++;
        W.cls_wt = p;
        off += align_up((size_t)D * Vp, 64); ++ti;
    }
    W.proj_w = plain(); W.proj_b = plain(); W.proj_ln_g = plain(); W.proj_ln_b = plain();
    if (rc == AC_OK && attn_emb_dim % 8 == 0) {
        const size_t np = align_up(tc_packed_floats(D, attn_emb_dim), 64), nk = align_up(tc_packed_floats(2 * D, D), 64);
        if (rc == AC_OK) rc = tc_pack_weight(W.proj_w, nullptr, D, attn_emb_dim, d->tcblob, st, &d->proj_tw);
        for (int l = 0; l < nlayers && rc == AC_OK; ++l)
            rc = tc_pack_weight(W.layer[l].ca_kv_w, nullptr, 2 * D, D, d->tcblob + np + l * nk, st, &d->kv_tw[l]);
    }
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "trm_fill pack kernels");
    return rc;
}

extern "C" {

int ac_trm_num_tensors(int nlayers) { return 2 + 18 * nlayers + 5; }

int ac_trm_create(const float* const* t, const int64_t* numels, int n_tensors, int d_model, int nhead, int nlayers,
                  int dim_ff, int vocab, int attn_emb_dim, int pe_len, void* stream, ac_trm_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_trm_create: null argument");
    AC_REQUIRE(d_model == D && nhead == NH, "ac_trm_create: only d_model=256 / nhead=4 is built (got %d/%d)", d_model, nhead);
    AC_REQUIRE(nlayers >= 1 && nlayers <= kMaxLayers, "ac_trm_create: nlayers %d not in [1,%d]", nlayers, kMaxLayers);
    AC_REQUIRE(dim_ff % 32 == 0 && dim_ff <= 1024, "ac_trm_create: dim_feedforward %d must be <=1024 and %%32", dim_ff);
    AC_REQUIRE(attn_emb_dim % 4 == 0, "ac_trm_create: attn_emb_dim %d must be a multiple of 4", attn_emb_dim);
    AC_REQUIRE(n_tensors == ac_trm_num_tensors(nlayers), "ac_trm_create: expected %d tensors, got %d",
               ac_trm_num_tensors(nlayers), n_tensors);
    cudaStream_t st = (cudaStream_t)stream;
    const int F = dim_ff, V = vocab;
    // expected element counts, in tensor order
    std::vector<int64_t> want = {(int64_t)V * D, (int64_t)pe_len * D};
    for (int l = 0; l < nlayers; ++l) {
        int64_t per[18] = {3 * D * D, 3 * D, D * D, D, 3 * D * D, 3 * D, D * D, D, (int64_t)F * D, F, (int64_t)D * F, D,
                           D, D, D, D, D, D};
        want.insert(want.end(), per, per + 18);
    }
    int64_t tail[5] = {(int64_t)V * D, (int64_t)D * attn_emb_dim, D, D, D};
    want.insert(want.end(), tail, tail + 5);
    for (int i = 0; i < n_tensors; ++i)
        AC_REQUIRE(numels[i] == want[i], "ac_trm_create: tensor %d has %lld elements, expected %lld", i,
                   (long long)numels[i], (long long)want[i]);
    size_t total = align_up((size_t)D * ((V + 3) / 4 * 4), 64);   // padded classifier copy
    for (auto n : want) total += align_up((size_t)n, 64);
    ac_trm_t* d = new ac_trm_t();
    d->want = want;
    DecW& W0 = d->w;
    W0.nlayers = nlayers; W0.dff = F; W0.vocab = V; W0.attn_emb_dim = attn_emb_dim; W0.pe_len = pe_len;
    int rc = check_cuda(cudaMalloc(&d->blob, total * sizeof(float)), "ac_trm_create: cudaMalloc");
    if (rc == AC_OK && attn_emb_dim % 8 == 0) {
        const size_t np = align_up(tc_packed_floats(D, attn_emb_dim), 64), nk = align_up(tc_packed_floats(2 * D, D), 64);
        rc = check_cuda(cudaMalloc(&d->tcblob, (np + nlayers * nk) * sizeof(float)), "cudaMalloc tc weights");
    }
    if (rc == AC_OK) rc = trm_fill(d, t, st);
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "ac_trm_create pack kernels");
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_trm_create sync");
    if (rc != AC_OK) { cudaFree(d->blob); cudaFree(d->tcblob); delete d; return rc; }
    *out = d;
    return AC_OK;
}

// Re-reads the source tensors (same order and sizes as at creation) into the existing handle: asynchronous, no allocation.
// The training loop calls it once per optimizer step so that the sampling decode of scheduled sampling sees the updated weights.
int ac_trm_update(ac_trm_t* d, const float* const* t, int n_tensors, void* stream) {
    using namespace ac;
    AC_REQUIRE(d && t && n_tensors == (int)d->want.size(), "ac_trm_update: bad argument");
    return trm_fill(d, t, (cudaStream_t)stream);
}

void ac_trm_destroy(ac_trm_t* d) {
    if (!d) return;
    cudaFree(d->blob);
    cudaFree(d->tcblob);
    delete d;
}

size_t ac_trm_workspace_bytes(const ac_trm_t* d, int rows, int t_mem, int max_len) {
    // `rows` = rows per clip (1 for greedy, beam size for beam search) is folded by the caller:
    // pass batch * rows_per_clip; sized for the worst case of both layouts.
    if (!d) return 0;
    ac::TrmWs s = ac::trm_ws(d, rows, 1, t_mem, max_len);
    return (s.proj + s.kvmem + s.cache + s.logits) * sizeof(float);
}

static long long* g_dec_trace = nullptr;

// `clusters` thread-block clusters of P CTAs
static int launch_decode(void (*kernel)(ac::DecodeArgs), int clusters, int P, size_t smem, cudaStream_t st,
                         const ac::DecodeArgs& a) {
    using namespace ac;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * P); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = P; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    AC_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
    return AC_OK;
}

static int trm_common_checks(const ac_trm_t* dec, int batch, int t_mem, int max_len) {
    using namespace ac;
    AC_REQUIRE(dec, "ac_trm: null decoder");
    AC_REQUIRE(batch >= 0, "ac_trm: negative batch");
    AC_REQUIRE(t_mem >= 1 && t_mem <= kMaxKeys, "ac_trm: t_mem %d not in [1,%d]", t_mem, kMaxKeys);
    AC_REQUIRE(max_len >= 1 && max_len <= kMaxLen && max_len <= dec->w.pe_len,
               "ac_trm: max_len %d not in [1,%d]", max_len, std::min(kMaxLen, dec->w.pe_len));
    return AC_OK;
}

static int trm_greedy_impl(const ac_trm_t* dec, const float* attn_emb, const int64_t* attn_emb_len, int batch, int t_mem,
                           int max_len, int start_idx, int end_idx, int pad_idx, const int64_t* forced, int train_mode,
                           int64_t* seq, float* logprob, float* logit, float* embed, void* ws, size_t ws_bytes, void* stream);

int ac_trm_greedy(const ac_trm_t* dec, const float* attn_emb, const int64_t* attn_emb_len, int batch, int t_mem,
                  int max_len, int start_idx, int end_idx, int pad_idx, int64_t* seq, float* logprob, float* logit,
                  float* embed, void* ws, size_t ws_bytes, void* stream) {
    return trm_greedy_impl(dec, attn_emb, attn_emb_len, batch, t_mem, max_len, start_idx, end_idx, pad_idx, nullptr, 0, seq,
                           logprob, logit, embed, ws, ws_bytes, stream);
}

// The sampling half of scheduled-sampling training (captioning/models/transformer_model.py:34-57 under
// captioning/models/base.py:152-170, mode == "train"): builds the model's own token row step by step, KV-cached.
// forced_dev [batch, max_len] int64: where >= 0 that token is emitted at the step (the step's coin chose the ground-truth
// prefix, whose arg-max the caller already knows from the dense pass), where < 0 the arg-max of this decode is taken.
// No early stop and no <end> forcing (the reference only does that in inference mode).  seq_dev [batch, max_len].
int ac_trm_sample_forced(const ac_trm_t* dec, const float* attn_emb, const int64_t* attn_emb_len, int batch, int t_mem,
                         int max_len, int start_idx, int end_idx, int pad_idx, const int64_t* forced_dev, int64_t* seq,
                         float* logprob, void* ws, size_t ws_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(forced_dev != nullptr, "ac_trm_sample_forced: forced_dev is NULL");
    return trm_greedy_impl(dec, attn_emb, attn_emb_len, batch, t_mem, max_len, start_idx, end_idx, pad_idx, forced_dev, 1, seq,
                           logprob, nullptr, nullptr, ws, ws_bytes, stream);
}

static int trm_greedy_impl(const ac_trm_t* dec, const float* attn_emb, const int64_t* attn_emb_len, int batch, int t_mem,
                           int max_len, int start_idx, int end_idx, int pad_idx, const int64_t* forced, int train_mode,
                           int64_t* seq, float* logprob, float* logit, float* embed, void* ws, size_t ws_bytes, void* stream) {
    using namespace ac;
    int rc = trm_common_checks(dec, batch, t_mem, max_len); if (rc) return rc;
    if (batch == 0) return AC_OK;
    AC_REQUIRE(attn_emb && attn_emb_len && seq, "ac_trm_greedy: null argument");
    TrmWs s = trm_ws(dec, batch, 1, t_mem, max_len);
    AC_REQUIRE(ws && ws_bytes >= (s.proj + s.kvmem + s.cache + s.logits) * sizeof(float), "ac_trm_greedy: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* proj = (float*)ws; float* kvmem = proj + s.proj; float* cache = kvmem + s.kvmem; float* lws = cache + s.cache;
    rc = prepare_memory(dec, attn_emb, batch, t_mem, proj, kvmem, st); if (rc) return rc;
    DecodeArgs a{};
    a.w = dec->w; a.kv_mem = kvmem; a.n_clips = batch; a.mem_len = attn_emb_len; a.kv_cache = cache; a.logits_ws = lws;
    a.t_mem = t_mem; a.max_len = max_len; a.start_idx = start_idx; a.end_idx = end_idx; a.pad_idx = pad_idx;
    a.seq = seq; a.logprob = logprob; a.logit_out = logit; a.embed_out = embed; a.beam = 1; a.temp = 1.0f;
    a.forced = forced; a.train_mode = train_mode;
    a.dbg = g_dec_trace;
    // Default: one CTA per attention head, G clips per 4-CTA cluster (greedy_heads_kernel) when its shared-memory plan fits;
    // AC_GREEDY="P,G" selects the column-split kernel below with that cluster shape (experiments, fallback).
    if (getenv("AC_GREEDY") == nullptr && dec->w.dff % (4 * NH) == 0 && dec->w.vocab >= 64) {
        int G = batch * NH <= kNumSMs ? 1 : 2;
        if (const char* e = getenv("AC_GREEDY_HEADS")) G = atoi(e);      // 0 disables
        if (G == 1 || G == 2 || G == 4) {
            const HeadSmem lay = head_smem(G, G, dec->w.nlayers, dec->w.dff, dec->w.vocab, t_mem, max_len);
            if (lay.total * sizeof(float) <= (size_t)kDecSmemLimit) {
                const size_t sm = lay.total * sizeof(float);
                AC_TIMED("trm_greedy", st);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cdiv(batch, G) * NH); cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = sm; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = NH; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
#define AC_HEADS_CASE(GG)                                                                                              \
    case GG:                                                                                                           \
        AC_CUDA(cudaFuncSetAttribute(greedy_heads_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
        AC_CUDA(cudaLaunchKernelEx(&cfg, greedy_heads_kernel<GG>, a, lay));                                            \
        break;
                switch (G) { AC_HEADS_CASE(1) AC_HEADS_CASE(2) AC_HEADS_CASE(4) }
#undef AC_HEADS_CASE
                AC_LAUNCHED("greedy_heads_kernel");
                return AC_OK;
            }
        }
    }
    // (CTAs per cluster, clips per cluster); AC_GREEDY="P,G" overrides for experiments
    int P = trm_cluster_size(batch, kGreedyCluster), G = kGreedyClips;
    if (const char* e = getenv("AC_GREEDY")) sscanf(e, "%d,%d", &P, &G);
    AC_REQUIRE((P == 1 || P == 2 || P == 4 || P == 8) && (G == 1 || G == 2 || G == 4), "ac_trm_greedy: bad cluster shape %d,%d", P, G);
    size_t sm = dec_smem_floats(G) * sizeof(float);
    const size_t kvb = kv_smem_floats(dec->w.nlayers, G, t_mem) * sizeof(float);
    a.kv_in_smem = sm + kvb <= (size_t)kDecSmemLimit ? 1 : 0;
    if (a.kv_in_smem) sm += kvb;
    AC_TIMED("trm_greedy", st);
#define AC_GREEDY_CASE(GG)                                                                                       \
    case GG:                                                                                                     \
        AC_CUDA(cudaFuncSetAttribute(greedy_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
        rc = launch_decode(greedy_kernel<GG>, cdiv(batch, GG), P, sm, st, a); if (rc) return rc;                 \
        break;
    switch (G) { AC_GREEDY_CASE(1) AC_GREEDY_CASE(2) AC_GREEDY_CASE(4) }
#undef AC_GREEDY_CASE
    AC_LAUNCHED("greedy_kernel");
    return AC_OK;
}

int ac_trm_beam(const ac_trm_t* dec, const float* attn_emb, const int64_t* attn_emb_len, int batch, int t_mem,
                int max_len, int beam, float temp, int start_idx, int end_idx, int pad_idx, int64_t* seq, void* ws,
                size_t ws_bytes, void* stream) {
    using namespace ac;
    int rc = trm_common_checks(dec, batch, t_mem, max_len); if (rc) return rc;
    if (batch == 0) return AC_OK;
    AC_REQUIRE(attn_emb && attn_emb_len && seq, "ac_trm_beam: null argument");
    AC_REQUIRE(beam >= 1 && beam <= 5, "ac_trm_beam: beam_size %d not in [1,5]", beam);
    AC_REQUIRE(temp > 0.f, "ac_trm_beam: temp must be > 0");
    if (batch == 0) return AC_OK;
    TrmWs s = trm_ws(dec, batch, beam, t_mem, max_len);
    AC_REQUIRE(ws && ws_bytes >= (s.proj + s.kvmem + s.cache + s.logits) * sizeof(float), "ac_trm_beam: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* proj = (float*)ws; float* kvmem = proj + s.proj; float* cache = kvmem + s.kvmem; float* lws = cache + s.cache;
    rc = prepare_memory(dec, attn_emb, batch, t_mem, proj, kvmem, st); if (rc) return rc;
    DecodeArgs a{};
    a.w = dec->w; a.kv_mem = kvmem; a.n_clips = batch; a.mem_len = attn_emb_len; a.kv_cache = cache; a.logits_ws = lws;
    a.t_mem = t_mem; a.max_len = max_len; a.start_idx = start_idx; a.end_idx = end_idx; a.pad_idx = pad_idx;
    a.seq = seq; a.beam = beam; a.temp = temp;
    // Up to 37 clips (4 CTAs per clip fit the SMs): one CTA per attention head (beam_heads_kernel), beam-3 2.38 -> 1.47 ms.
    // Beyond, the column-split kernel with 2 CTAs per clip stays faster than two waves of 4-CTA clusters (2.76 vs 2.94 ms at
    // 64 clips) and than two clips x three beams per cluster (3.0-3.5 ms: six rows leave the split-K scratch room for a
    // third of the threads only).  AC_BEAM_HEADS=0 / 1 forces the choice.
    {
        const char* e = getenv("AC_BEAM_HEADS");
        const bool want = e ? atoi(e) != 0 : batch * NH <= kNumSMs;
        if (want && dec->w.dff % (4 * NH) == 0 && dec->w.vocab >= 64) {
            HeadSmem lay = head_smem(beam, 1, dec->w.nlayers, dec->w.dff, dec->w.vocab, t_mem, max_len);
            if (lay.total * sizeof(float) > (size_t)kDecSmemLimit &&
                (size_t)batch * NH * dec->w.nlayers * beam * max_len * 2 * HD <= s.cache)
                lay = head_smem(beam, 1, dec->w.nlayers, dec->w.dff, dec->w.vocab, t_mem, max_len, true);
            if (lay.total * sizeof(float) <= (size_t)kDecSmemLimit) {
                const size_t hsm = lay.total * sizeof(float);
                AC_TIMED("trm_beam", st);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(batch * NH); cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = hsm; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = NH; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
#define AC_BH(RR)                                                                                                      \
    case RR:                                                                                                           \
        AC_CUDA(cudaFuncSetAttribute(beam_heads_kernel<RR, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm)); \
        AC_CUDA(cudaLaunchKernelEx(&cfg, beam_heads_kernel<RR, 1>, a, lay));                                           \
        break;
                switch (beam) { AC_BH(1) AC_BH(2) AC_BH(3) AC_BH(4) AC_BH(5) }
#undef AC_BH
                AC_LAUNCHED("beam_heads_kernel");
                return AC_OK;
            }
        }
    }
    size_t sm = dec_smem_floats(beam) * sizeof(float);
    const size_t kvb = kv_smem_floats(dec->w.nlayers, 1, t_mem) * sizeof(float);
    a.kv_in_smem = sm + kvb <= (size_t)kDecSmemLimit ? 1 : 0;
    if (a.kv_in_smem) sm += kvb;
    AC_TIMED("trm_beam", st);
#define AC_BEAM_CASE(RR)                                                                                        \
    case RR:                                                                                                    \
        AC_CUDA(cudaFuncSetAttribute(beam_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   \
        rc = launch_decode(beam_kernel<RR>, batch, trm_cluster_size(batch, kBeamCluster), sm, st, a); if (rc) return rc; \
        break;
    switch (beam) {
        AC_BEAM_CASE(1) AC_BEAM_CASE(2) AC_BEAM_CASE(3) AC_BEAM_CASE(4) AC_BEAM_CASE(5)
    }
#undef AC_BEAM_CASE
    AC_LAUNCHED("beam_kernel");
    return AC_OK;
}

/* Diagnostic: phase trace (clock64) of decode step 5 of CTA 0 of the next greedy launches; out_host: 64 int64. */
int ac_trm_trace(int on, long long* out_host) {
    using namespace ac;
    if (on && g_dec_trace == nullptr) {
        AC_CUDA(cudaMalloc(&g_dec_trace, 64 * sizeof(long long)));
        AC_CUDA(cudaMemset(g_dec_trace, 0, 64 * sizeof(long long)));
    }
    if (out_host != nullptr && g_dec_trace != nullptr) {
        AC_CUDA(cudaDeviceSynchronize());
        AC_CUDA(cudaMemcpy(out_host, g_dec_trace, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    if (!on && g_dec_trace != nullptr) { cudaFree(g_dec_trace); g_dec_trace = nullptr; }
    return AC_OK;
}

}  // extern "C"
