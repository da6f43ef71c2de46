// Temporal Bahdanau-attention GRU caption decoder: greedy and beam-search decoding in one launch, fp32.
//
// Replaces captioning/models/hf_wrapper.py:1377-1414 `Seq2SeqAttention`, :1444-1554 `BahAttnCatFcDecoder` /
// `TemporalBahAttnDecoder.forward` and the decode glue :1557-1788 (`Seq2SeqAttnModel`, `TemporalSeq2SeqAttnModel`)
// driven by `CaptionModel.stepwise_forward` / `beam_search` (captioning/models/base.py:152-218, 254-361), eval mode.
//
// Per step and row (h = GRU state, enc = audio memory [T, 512]):
//   score_s = v . tanh(W_q h + (W_e enc_s + b)),  masked past the clip's length, softmax, ctx = sum_s w_s enc_s
//   x = [embed ; ctx_proj(ctx) ; fc_proj(fc_emb)],  h' = GRUCell(x, h),  logit = classifier(h') (+ bias)
// Work that does not depend on the step is hoisted out of it:
//   * E = enc W_e^T + b (per call, one tensor-core GEMM) -- the reference recomputes it for every step;
//   * W_ih[:, 0:512] . embedding is a table lookup (packed once per weight load, [V + 4 tags][1536]);
//   * W_ih[:, 512:1024] . ctx_proj folds into ONE [1536 x 512] matrix applied to ctx;
//   * W_ih[:, 1024:1536] . fc_proj(fc_emb) + all input-side biases is one vector per clip (two small GEMMs per call).
// What remains per step is three GEMVs (W_q h, W_hh h, M_c ctx: 7.3 MB of weights) + the classifier (10.2 MB), streamed
// from L2 by a 1/2/4-CTA cluster per clip (chosen from the batch size) exactly like the Transformer decoder (decode_common.cuh: each CTA computes half
// of the output columns and writes them into both CTAs' shared memory); attention, the cell update and the
// arg-max / top-k bookkeeping run redundantly in both CTAs.  Beam search: the R beams of a clip are the R rows of one
// cluster (every weight load is shared by R rows); the GRU state follows `prev_words_beam` (hf_wrapper.py:1656-1661).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "decode_common.cuh"
#include "gemm.cuh"

namespace ac {

constexpr int BH = 512;            // embedding = hidden = attention = memory = fc width of the released model
constexpr int BG = 3 * BH;
constexpr int kBahMaxT = 128;      // memory frames (41 s of 32 kHz audio; s_sc is only R x kBahMaxT floats)
constexpr int kBahMaxLen = 64;     // decode steps
// CTAs per clip: small batches spread each clip's 17.5 MB of weights per step over more SMs, large ones keep one wave
static int bah_cluster_size(int clips) {
    if (const char* e = getenv("AC_BAH_CLUSTER")) { const int p = atoi(e); if (p == 1 || p == 2 || p == 4 || p == 8) return p; }   // tuning aid
    // measured on B200 (scripts/bah_time.py): 16 clips beam-4 2.2 ms at P = 4 vs 3.2 ms at P = 2 / 3.3 ms at P = 8;
    // 128 clips 5.5 ms at P = 1 vs 6.3 ms at P = 2 (a second wave of CTAs costs more than the wider split saves)
    return clips <= kNumSMs / 4 ? 4 : clips <= kNumSMs / 2 ? 2 : 1;
}

struct BahW {
    const float* emb_tab;   // [V + 4][BG]  W_ih[:, 0:512] . (word | temporal) embedding
    const float* mc_t;      // [BH][BG]     (W_ih[:, 512:1024] . W_ctx)^T
    const float* whh_t;     // [BH][BG]
    const float* bhh;       // [BG]
    const float* wq_t;      // [BH][BH]     h2attn.weight[:, 0:512]^T
    const float* v;         // [BH]
    const float* cls_t;     // [BH][Vp]
    const float* cls_b;     // [Vp]
    int vocab, vp;
};

struct BahArgs {
    BahW w;
    const float* E;           // [clips][T][BH]
    const float* enc;         // [clips][T][BH]
    const float* gfc;         // [clips][BG]
    const int64_t* mem_len;   // [clips]
    const int64_t* tags;      // [clips] in 0..3
    float* logits_ws;         // [clips][R][Vp]
    int n_clips, T, max_len, start_idx, end_idx;
    int64_t* seq;             // [clips][max_len]
    float* logprob;           // nullable [clips][max_len]
    float* logit_out;         // nullable [clips][max_len][V]
    int beam; float temp;
    // single-step / stepwise extras of the greedy kernel (`TemporalBahAttnDecoder.forward`, hf_wrapper.py:1513-1554)
    const float* h0;          // nullable [clips][BH]: initial GRU state (default zeros, `init_hidden`)
    float* state_out;         // nullable [clips][BH]: GRU state after the last executed step
    float* attn_w_out;        // nullable [clips][max_len][T]: attention weights of every step
    const int64_t* first_tok; // nullable [clips]: word fed at step 0 instead of the temporal-tag embedding
};

// One decoder step for the R rows of this cluster.  s_tok[r] = row index into emb_tab; s_h is updated in place.
template <int R>
__device__ __forceinline__ void bah_step(const BahArgs& a, int clip, const int* s_tok, float* s_h, float* s_hn, float* s_q,
                                         float* s_ctx, float* s_gh, float* s_gc, float* s_sc, float* s_part,
                                         float* const* logits, float* aw = nullptr) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), P = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BahW& W = a.w;
    const int T = a.T;
    const int len = (int)min((int64_t)T, max((int64_t)1, a.mem_len[clip]));
    // ---- state-side GEMVs
    matvec_t<R>(W.wq_t, nullptr, s_h, BH, s_q, BH, BH, BH, false, s_part, rank, P);
    cluster.sync();
    matvec_t<R>(W.whh_t, W.bhh, s_h, BH, s_gh, BG, BG, BH, false, s_part, rank, P);
    cluster.sync();
    // ---- additive attention: one warp per (row, frame)
    const float* E = a.E + (size_t)clip * T * BH;
    const float* enc = a.enc + (size_t)clip * T * BH;
    for (int i = warp; i < R * T; i += kWarps) {
        const int r = i / T, s = i - r * T;
        float acc = 0.f;
        if (s < len) {
            const float* e = E + (size_t)s * BH;
            const float* q = s_q + r * BH;
            float ev[BH / 32];
#pragma unroll
            for (int j = 0; j < BH / 32; ++j) ev[j] = __ldg(e + lane + 32 * j);      // all 16 L2 loads in flight at once
#pragma unroll
            for (int j = 0; j < BH / 32; ++j) acc = fmaf(__ldg(W.v + lane + 32 * j), tanhf(q[lane + 32 * j] + ev[j]), acc);
            acc = warp_sum(acc);
        }
        if (lane == 0) s_sc[r * kBahMaxT + s] = s < len ? acc : -1e10f;   // masked_fill(mask == 0, -1e10)
    }
    __syncthreads();
    if (warp < R) {
        float* sc = s_sc + warp * kBahMaxT;
        float m = -INFINITY;
        for (int s = lane; s < T; s += 32) m = fmaxf(m, sc[s]);
        m = warp_max(m);
        float se = 0.f;
        for (int s = lane; s < T; s += 32) { const float e = expf(sc[s] - m); sc[s] = e; se += e; }
        se = warp_sum(se);
        const float inv = 1.0f / se;
        for (int s = lane; s < T; s += 32) {
            sc[s] *= inv;
            if (aw != nullptr && rank == 0) aw[warp * T + s] = sc[s];       // `attn_weight` of the step (row-major [R][T])
        }
    }
    __syncthreads();
    for (int i = tid; i < R * BH; i += kThreads) {
        const int r = i / BH, c = i - r * BH;
        const float* sc = s_sc + r * kBahMaxT;
        float acc = 0.f;
        for (int s0 = 0; s0 < len; s0 += 8) {                  // 8 independent L2 loads per batch
            float ev[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ev[j] = s0 + j < len ? __ldg(enc + (size_t)(s0 + j) * BH + c) : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(s0 + j < len ? sc[s0 + j] : 0.0f, ev[j], acc);
        }
        s_ctx[i] = acc;
    }
    __syncthreads();
    // ---- context-side GEMV and the GRU cell
    matvec_t<R>(W.mc_t, nullptr, s_ctx, BH, s_gc, BG, BG, BH, false, s_part, rank, P);
    cluster.sync();
    const float* gf = a.gfc + (size_t)clip * BG;
    for (int i = tid; i < R * BH; i += kThreads) {
        const int r = i / BH, u = i - r * BH;
        const float* te = W.emb_tab + (size_t)s_tok[r] * BG;
        const float* gc = s_gc + r * BG;
        const float* gh = s_gh + r * BG;
        const float gi_r = __ldg(te + u) + gc[u] + __ldg(gf + u);
        const float gi_z = __ldg(te + BH + u) + gc[BH + u] + __ldg(gf + BH + u);
        const float gi_n = __ldg(te + 2 * BH + u) + gc[2 * BH + u] + __ldg(gf + 2 * BH + u);
        const float rg = 1.0f / (1.0f + expf(-(gi_r + gh[u])));
        const float zg = 1.0f / (1.0f + expf(-(gi_z + gh[BH + u])));
        const float ng = tanhf(gi_n + rg * gh[2 * BH + u]);
        s_hn[i] = (1.0f - zg) * ng + zg * s_h[i];
    }
    __syncthreads();
    for (int i = tid; i < R * BH; i += kThreads) s_h[i] = s_hn[i];
    __syncthreads();
    // ---- classifier (with bias): each thread owns column quads of the zero-padded [BH][Vp] weight; when the CTA's
    // share of the columns is smaller than the block (large clusters) the k range is split across thread groups
    const int V = W.vocab, VC = W.vp >> 2;
    const int vc0 = (int)((int64_t)VC * rank / P), vc1 = (int)((int64_t)VC * (rank + 1) / P);
    const int ncq = vc1 - vc0;
    const int KS = ncq >= kThreads ? 1 : max(1, min(kThreads / ncq, 4096 / (4 * ncq)));
    const int kslice = (BH + KS - 1) / KS;
    for (int cc = tid; cc < ncq * KS; cc += kThreads) {
        const int ks = cc / ncq, c = vc0 + (cc - ks * ncq);
        const int k0 = ks * kslice, k1 = min(BH, k0 + kslice);
        float acc[R][4];
        const float4 b4 = ks == 0 ? __ldg(reinterpret_cast<const float4*>(W.cls_b) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = b4.x; acc[r][1] = b4.y; acc[r][2] = b4.z; acc[r][3] = b4.w; }
        const float4* w = reinterpret_cast<const float4*>(W.cls_t) + c;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) {
            const float4 wv = __ldg(w + (size_t)k * VC);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float xv = s_h[r * BH + k];
                acc[r][0] = fmaf(wv.x, xv, acc[r][0]); acc[r][1] = fmaf(wv.y, xv, acc[r][1]);
                acc[r][2] = fmaf(wv.z, xv, acc[r][2]); acc[r][3] = fmaf(wv.w, xv, acc[r][3]);
            }
        }
        if (KS == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * c + q < V) logits[r][4 * c + q] = acc[r][q];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                *reinterpret_cast<float4*>(s_part + ((size_t)(ks * R + r) * ncq + (c - vc0)) * 4) =
                    make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    }
    if (KS > 1) {
        __syncthreads();
        for (int i = tid; i < R * ncq * 4; i += kThreads) {
            const int r = i / (ncq * 4), j = i - r * (ncq * 4);
            float v = 0.f;
            for (int ks = 0; ks < KS; ++ks) v += s_part[(size_t)(ks * R + r) * ncq * 4 + j];   // fixed order
            const int n = 4 * vc0 + j;
            if (n < V) logits[r][n] = v;
        }
    }
    cluster.sync();   // both halves of the logits (global memory) are visible to both CTAs
}

#define BAH_SMEM_LAYOUT(R)                                                                                       \
    extern __shared__ __align__(16) float smem[];                                                               \
    float* s_h = smem; float* s_hn = s_h + R * BH; float* s_q = s_hn + R * BH; float* s_ctx = s_q + R * BH;     \
    float* s_gh = s_ctx + R * BH; float* s_gc = s_gh + R * BG; float* s_sc = s_gc + R * BG;                     \
    float* s_part = s_sc + R * kBahMaxT;
static size_t bah_smem_bytes(int R) { return (size_t)R * (4 * BH + 2 * BG + kBahMaxT + 4096) * sizeof(float); }

// ------------------------------------------------------------------------------------ greedy (one clip per cluster)
__global__ void __launch_bounds__(kThreads)
bah_greedy_kernel(BahArgs a) {
    BAH_SMEM_LAYOUT(1)
    __shared__ float s_rv[kWarps];
    __shared__ int s_ri[kWarps];
    __shared__ int s_tok[1];
    __shared__ float* s_logits[1];
    cg::cluster_group cluster = cg::this_cluster();
    const int P = (int)cluster.num_blocks();
    const int clip = blockIdx.x / P, tid = threadIdx.x;
    const bool writer = cluster.block_rank() == 0;
    const int V = a.w.vocab;
    for (int i = tid; i < BH; i += kThreads) s_h[i] = a.h0 ? a.h0[(size_t)clip * BH + i] : 0.0f;      // init_hidden: zeros
    int word = a.start_idx;
    bool finished = false;
    cluster.sync();   // every CTA of the cluster is resident before the first GEMV writes into its peers' shared memory
    // like the reference, a finished row keeps running (input forced to <end>) while logits are requested
    const bool full_outputs = a.logit_out != nullptr;
    for (int t = 0; t < a.max_len; ++t) {
        if (finished && !full_outputs) {
            if (tid == 0 && writer) a.seq[(size_t)clip * a.max_len + t] = a.end_idx;
            continue;
        }
        if (tid == 0) {
            s_tok[0] = t > 0 ? word
                     : a.first_tok ? (int)min((int64_t)V - 1, max((int64_t)0, a.first_tok[clip]))
                                   : V + (int)min((int64_t)3, max((int64_t)0, a.tags[clip]));
            s_logits[0] = a.logit_out ? a.logit_out + ((size_t)clip * a.max_len + t) * V : a.logits_ws + (size_t)clip * a.w.vp;
        }
        __syncthreads();
        bah_step<1>(a, clip, s_tok, s_h, s_hn, s_q, s_ctx, s_gh, s_gc, s_sc, s_part, s_logits,
                    a.attn_w_out ? a.attn_w_out + ((size_t)clip * a.max_len + t) * a.T : nullptr);
        if (a.state_out != nullptr && writer)
            for (int i = tid; i < BH; i += kThreads) a.state_out[(size_t)clip * BH + i] = s_h[i];
        const float* logits = s_logits[0];
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int n = tid; n < V; n += kThreads) {
            const float v = logits[n];
            if (v > best) { best = v; bi = n; }
        }
        block_argmax(best, bi, s_rv, s_ri);
        float se = 0.f;
        for (int n = tid; n < V; n += kThreads) se += expf(logits[n] - best);
        se = block_sum(se, s_rv);
        word = finished ? a.end_idx : bi;
        if (tid == 0 && writer) {
            a.seq[(size_t)clip * a.max_len + t] = word;
            if (a.logprob) a.logprob[(size_t)clip * a.max_len + t] = -logf(se);
        }
        finished = finished || (word == a.end_idx);
    }
}

// ------------------------------------------------------------------------------------ beam search (R beams of one clip)
template <int R>
__global__ void __launch_bounds__(kThreads)
bah_beam_kernel(BahArgs a) {
    BAH_SMEM_LAYOUT(R)
    __shared__ float* s_logits[R];
    __shared__ int s_seq[2][R][kBahMaxLen];
    __shared__ int s_tok[R];
    __shared__ int s_parent[R];
    __shared__ float s_score[R];
    __shared__ float s_newscore[R];
    __shared__ int s_newidx[R];
    __shared__ float s_rv[kWarps];
    __shared__ int s_ri[kWarps];
    __shared__ int s_best_seq[kBahMaxLen];
    __shared__ int s_best_len, s_ndone, s_stop;
    __shared__ float s_best_score;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), P = (int)cluster.num_blocks();
    const int clip = blockIdx.x / P, tid = threadIdx.x;
    const int V = a.w.vocab, Vp = a.w.vp;
    float* lp = a.logits_ws + (size_t)clip * R * Vp;
    const int tag_tok = V + (int)min((int64_t)3, max((int64_t)0, a.tags[clip]));
    if (tid < R) { s_tok[tid] = tag_tok; s_score[tid] = 0.f; s_logits[tid] = lp + (size_t)tid * Vp; s_parent[tid] = tid; }
    if (tid == 0) { s_ndone = 0; s_stop = 0; s_best_len = 0; s_best_score = -INFINITY; }
    for (int i = tid; i < R * BH; i += kThreads) s_h[i] = 0.0f;
    cluster.sync();   // peers resident before the first distributed-shared-memory write (decode_common.cuh matvec_t)
    int cur = 0;
    for (int t = 0; t < a.max_len; ++t) {
        if (t > 0) {   // the state follows its parent beam (state[:, prev_words_beam, :])
            for (int i = tid; i < R * BH; i += kThreads) s_hn[i] = s_h[s_parent[i / BH] * BH + (i % BH)];
            __syncthreads();
            for (int i = tid; i < R * BH; i += kThreads) s_h[i] = s_hn[i];
            __syncthreads();
        }
        bah_step<R>(a, clip, s_tok, s_h, s_hn, s_q, s_ctx, s_gh, s_gc, s_sc, s_part, s_logits);
        // lp = log_softmax(log_softmax(logit) / temp) + running score   (base.py:282-290)
        for (int r = 0; r < R; ++r) {
            float* row = lp + (size_t)r * Vp;
            float m = -INFINITY; int mi = 0;
            for (int n = tid; n < V; n += kThreads) m = fmaxf(m, row[n]);
            block_argmax(m, mi, s_rv, s_ri);
            float se = 0.f;
            for (int n = tid; n < V; n += kThreads) se += expf(row[n] - m);
            const float lse1 = m + logf(block_sum(se, s_rv));
            const float inv_t = 1.0f / a.temp;
            const float m2 = (m - lse1) * inv_t;
            float se2 = 0.f;
            for (int n = tid; n < V; n += kThreads) se2 += expf((row[n] - lse1) * inv_t - m2);
            const float lse2 = m2 + logf(block_sum(se2, s_rv));
            const float sc = s_score[r];
            const int n0 = (int)((int64_t)V * rank / P), n1 = (int)((int64_t)V * (rank + 1) / P);
            cluster.sync();   // every read of the raw logits above is done in both CTAs
            for (int n = n0 + tid; n < n1; n += kThreads) row[n] = sc + ((row[n] - lse1) * inv_t - lse2);
        }
        cluster.sync();
        // top-R of the flattened [rows][V] scores (step 0: row 0 only); flat index = row * V + word as in the reference
        const int nrows = t == 0 ? 1 : R;
        for (int k = 0; k < R; ++k) {
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int rr = 0; rr < nrows; ++rr)
                for (int n = tid; n < V; n += kThreads) {
                    const int flat = rr * V + n;
                    bool taken = false;
                    for (int q = 0; q < k; ++q) taken |= (s_newidx[q] == flat);
                    const float v = lp[(size_t)rr * Vp + n];
                    if (!taken && (v > best || (v == best && flat < bi))) { best = v; bi = flat; }
                }
            block_argmax(best, bi, s_rv, s_ri);
            if (tid == 0) { s_newscore[k] = best; s_newidx[k] = bi; }
            __syncthreads();
        }
        const int nxt = cur ^ 1;
        for (int i = tid; i < R * kBahMaxLen; i += kThreads) {
            const int r = i / kBahMaxLen, j = i % kBahMaxLen;
            const int parent = s_newidx[r] / V;
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
                s_score[r] = sc; s_tok[r] = wd; s_parent[r] = s_newidx[r] / V;
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

// ------------------------------------------------------------------------------------ pack helpers
// out[r][c] = in[r * ld + col0 + c]
__global__ void bah_slice_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int ld, int col0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)rows * cols) out[i] = in[(i / cols) * ld + col0 + (i % cols)];
}
// in [rows][cols] -> out [cols][ld] (ld >= rows; the padding stays zero)
__global__ void bah_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)rows * cols) out[(i % cols) * ld + i / cols] = in[i];
}

}  // namespace ac

struct ac_bah {
    float* blob = nullptr;
    ac::BahW w;
    const float *we, *b_att, *w_fc, *b_fc, *w_ihf, *cbias;   // per-call GEMM operands (original [N][K] layout)
    ac::TcWeight we_tw, wfc_tw, wihf_tw;
};

extern "C" {

int ac_bah_num_tensors(void) { return 15; }

int ac_bah_create(const float* const* t, const int64_t* numels, int n_tensors, int vocab, void* stream, ac_bah_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_bah_create: null argument");
    AC_REQUIRE(n_tensors == 15, "ac_bah_create: expected 15 tensors, got %d", n_tensors);
    AC_REQUIRE(vocab >= 4, "ac_bah_create: bad vocabulary size %d", vocab);
    const int V = vocab, Vp = (V + 3) / 4 * 4;
    const int64_t want[15] = {(int64_t)V * BH, (int64_t)V * BH, V, (int64_t)BG * BG, (int64_t)BG * BH, BG, BG, BH,
                              (int64_t)BH * 2 * BH, BH, (int64_t)BH * BH, BH, (int64_t)BH * BH, BH, 4 * BH};
    for (int i = 0; i < 15; ++i)
        AC_REQUIRE(numels[i] == want[i], "ac_bah_create: tensor %d has %lld elements, expected %lld (the kernel is built "
                   "for emb = hidden = attention = memory = fc width 512)", i, (long long)numels[i], (long long)want[i]);
    const float *emb = t[0], *cls_w = t[1], *cls_b = t[2], *w_ih = t[3], *w_hh = t[4], *b_ih = t[5], *b_hh = t[6], *v = t[7],
                *w_att = t[8], *b_att = t[9], *w_fc = t[10], *b_fc = t[11], *w_ctx = t[12], *b_ctx = t[13], *temb = t[14];
    cudaStream_t st = (cudaStream_t)stream;
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += align_up(n, 64); return o; };
    const size_t o_tab = take((size_t)(V + 4) * BG), o_mct = take((size_t)BH * BG), o_whht = take((size_t)BH * BG), o_bhh = take(BG),
                 o_wqt = take((size_t)BH * BH), o_v = take(BH), o_clst = take((size_t)BH * Vp), o_clsb = take(Vp),
                 o_we = take((size_t)BH * BH), o_batt = take(BH), o_wfc = take((size_t)BH * BH), o_bfc = take(BH),
                 o_wihf = take((size_t)BG * BH), o_cbias = take(BG), o_we_pk = take(tc_packed_floats(BH, BH)),
                 o_wfc_pk = take(tc_packed_floats(BH, BH)), o_wihf_pk = take(tc_packed_floats(BG, BH));
    // scratch: all embeddings [V+4][BH], W_ih[:, 0:512], W_ih[:, 512:1024], W_ctx^T
    const size_t s_x = take((size_t)(V + 4) * BH), s_wihe = take((size_t)BG * BH), s_wihc = take((size_t)BG * BH),
                 s_wctxt = take((size_t)BH * BH);
    ac_bah_t* d = new ac_bah_t();
    int rc = check_cuda(cudaMalloc(&d->blob, total * sizeof(float)), "ac_bah_create: cudaMalloc");
    if (rc != AC_OK) { delete d; return rc; }
    float* B0 = d->blob;
    rc = check_cuda(cudaMemsetAsync(B0, 0, total * sizeof(float), st), "memset");
    auto copy = [&](size_t off, const float* src, size_t n) {
        if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(B0 + off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy");
    };
    auto slice = [&](size_t off, const float* src, int rows, int cols, int ld, int col0) {
        bah_slice_kernel<<<(unsigned)cdiv64((int64_t)rows * cols, 256), 256, 0, st>>>(src, B0 + off, rows, cols, ld, col0);
        g_launches++;
    };
    auto transpose = [&](size_t off, const float* src, int rows, int cols, int ld) {
        bah_transpose_kernel<<<(unsigned)cdiv64((int64_t)rows * cols, 256), 256, 0, st>>>(src, B0 + off, rows, cols, ld);
        g_launches++;
    };
    auto simt = [&](const float* A, const float* Wm, float* C, int M, int N, int K, const float* bias) {
        if (rc != AC_OK) return;
        GemmArgs g; g.A = A; g.W = Wm; g.C = C; g.M = M; g.N = N; g.K = K; g.cbias = bias; g.act = ACT_NONE;
        rc = gemm_tn_simt(g, st);
    };
    copy(s_x, emb, (size_t)V * BH);
    copy(s_x + (size_t)V * BH, temb, 4 * BH);
    slice(s_wihe, w_ih, BG, BH, BG, 0);
    slice(s_wihc, w_ih, BG, BH, BG, BH);
    slice(o_wihf, w_ih, BG, BH, BG, 2 * BH);
    transpose(s_wctxt, w_ctx, BH, BH, BH);
    simt(B0 + s_x, B0 + s_wihe, B0 + o_tab, V + 4, BG, BH, nullptr);            // emb_tab = X . W_ih_e^T
    simt(B0 + s_wctxt, B0 + s_wihc, B0 + o_mct, BH, BG, BH, nullptr);           // mc_t[k][n] = sum_j W_ctx[j][k] W_ih_c[n][j]
    simt(b_ctx, B0 + s_wihc, B0 + o_cbias, 1, BG, BH, b_ih);                    // cbias = W_ih_c b_ctx + b_ih
    transpose(o_whht, w_hh, BG, BH, BG);
    copy(o_bhh, b_hh, BG);
    slice(o_we, w_att, BH, BH, 2 * BH, BH);                                     // W_e = h2attn.weight[:, 512:1024]
    {   // W_q^T: slice then transpose through the (now free) W_ctx^T scratch is not possible before the GEMM above ran
        // in stream order it is: all launches are on `st`
        slice(s_wihe, w_att, BH, BH, 2 * BH, 0);
        transpose(o_wqt, B0 + s_wihe, BH, BH, BH);
    }
    copy(o_v, v, BH);
    copy(o_batt, b_att, BH);
    transpose(o_clst, cls_w, V, BH, Vp);
    copy(o_clsb, cls_b, V);
    copy(o_wfc, w_fc, (size_t)BH * BH);
    copy(o_bfc, b_fc, BH);
    if (rc == AC_OK) rc = tc_pack_weight(B0 + o_we, nullptr, BH, BH, B0 + o_we_pk, st, &d->we_tw);
    if (rc == AC_OK) rc = tc_pack_weight(B0 + o_wfc, nullptr, BH, BH, B0 + o_wfc_pk, st, &d->wfc_tw);
    if (rc == AC_OK) rc = tc_pack_weight(B0 + o_wihf, nullptr, BG, BH, B0 + o_wihf_pk, st, &d->wihf_tw);
    if (rc == AC_OK) rc = check_cuda(cudaGetLastError(), "ac_bah_create pack kernels");
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_bah_create sync");
    if (rc != AC_OK) { cudaFree(d->blob); delete d; return rc; }
    BahW& W = d->w;
    W.emb_tab = B0 + o_tab; W.mc_t = B0 + o_mct; W.whh_t = B0 + o_whht; W.bhh = B0 + o_bhh; W.wq_t = B0 + o_wqt; W.v = B0 + o_v;
    W.cls_t = B0 + o_clst; W.cls_b = B0 + o_clsb; W.vocab = V; W.vp = Vp;
    d->we = B0 + o_we; d->b_att = B0 + o_batt; d->w_fc = B0 + o_wfc; d->b_fc = B0 + o_bfc; d->w_ihf = B0 + o_wihf; d->cbias = B0 + o_cbias;
    *out = d;
    return AC_OK;
}

void ac_bah_destroy(ac_bah_t* d) {
    if (!d) return;
    cudaFree(d->blob);
    delete d;
}

size_t ac_bah_workspace_bytes(const ac_bah_t* d, int rows, int T) {
    if (!d) return 0;
    using namespace ac;
    return (align_up((size_t)rows * T * BH, 64) + align_up((size_t)rows * BH, 64) + align_up((size_t)rows * BG, 64) +
            align_up((size_t)rows * d->w.vp, 64)) * sizeof(float);
}

}  // extern "C"

namespace ac {
// per-call hoisted work: E = enc W_e^T + b_att,  gfc = W_ih_f (W_fc fc + b_fc) + cbias
static int bah_prepare(const ac_bah_t* d, const float* fc_emb, const float* attn_emb, int B, int T, float* ws, int rows,
                       BahArgs& a, cudaStream_t st) {
    float* E = ws;
    float* P1 = E + align_up((size_t)rows * T * BH, 64);
    float* gfc = P1 + align_up((size_t)rows * BH, 64);
    float* logits = gfc + align_up((size_t)rows * BG, 64);
    GemmArgs g; g.A = attn_emb; g.W = d->we; g.C = E; g.M = B * T; g.N = BH; g.K = BH; g.cbias = d->b_att; g.act = ACT_NONE; g.tw = &d->we_tw;
    int rc = gemm_tn(g, st); if (rc) return rc;
    GemmArgs g1; g1.A = fc_emb; g1.W = d->w_fc; g1.C = P1; g1.M = B; g1.N = BH; g1.K = BH; g1.cbias = d->b_fc; g1.act = ACT_NONE; g1.tw = &d->wfc_tw;
    rc = gemm_tn(g1, st); if (rc) return rc;
    GemmArgs g2; g2.A = P1; g2.W = d->w_ihf; g2.C = gfc; g2.M = B; g2.N = BG; g2.K = BH; g2.cbias = d->cbias; g2.act = ACT_NONE; g2.tw = &d->wihf_tw;
    rc = gemm_tn(g2, st); if (rc) return rc;
    a.w = d->w; a.E = E; a.enc = attn_emb; a.gfc = gfc; a.logits_ws = logits; a.n_clips = B; a.T = T;
    return AC_OK;
}
template <typename K>
static int bah_launch(K kernel, int clusters, size_t smem, cudaStream_t st, const BahArgs& a) {
    AC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int kBahCluster = bah_cluster_size(clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kBahCluster); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kBahCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AC_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
    return AC_OK;
}
}  // namespace ac

extern "C" {

int ac_bah_greedy(const ac_bah_t* d, const float* fc_emb, const float* attn_emb, const int64_t* lens, const int64_t* tags,
                  int B, int T, int max_len, int start_idx, int end_idx, int64_t* seq, float* logprob, float* logit_out,
                  void* ws, size_t ws_bytes, void* stream) {
    return ac_bah_greedy_ex(d, fc_emb, attn_emb, lens, tags, B, T, max_len, start_idx, end_idx, nullptr, nullptr, seq, logprob,
                            logit_out, nullptr, nullptr, ws, ws_bytes, stream);
}

int ac_bah_greedy_ex(const ac_bah_t* d, const float* fc_emb, const float* attn_emb, const int64_t* lens, const int64_t* tags,
                     int B, int T, int max_len, int start_idx, int end_idx, const float* state_in, const int64_t* first_word,
                     int64_t* seq, float* logprob, float* logit_out, float* state_out, float* attn_w_out,
                     void* ws, size_t ws_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 0 && T >= 1 && T <= kBahMaxT && max_len >= 1 && max_len <= kBahMaxLen,
               "ac_bah_greedy: sizes out of range (B=%d T=%d max_len=%d; T <= %d, max_len <= %d)", B, T, max_len, kBahMaxT, kBahMaxLen);
    if (B == 0) return AC_OK;
    AC_REQUIRE(d && fc_emb && attn_emb && lens && tags && seq, "ac_bah_greedy: null argument");
    AC_REQUIRE(ws && ws_bytes >= ac_bah_workspace_bytes(d, B, T), "ac_bah_greedy: workspace too small (%zu < %zu)", ws_bytes,
               ac_bah_workspace_bytes(d, B, T));
    cudaStream_t st = (cudaStream_t)stream;
    BahArgs a{};
    int rc = bah_prepare(d, fc_emb, attn_emb, B, T, (float*)ws, B, a, st); if (rc) return rc;
    a.mem_len = lens; a.tags = tags; a.max_len = max_len; a.start_idx = start_idx; a.end_idx = end_idx;
    a.seq = seq; a.logprob = logprob; a.logit_out = logit_out; a.beam = 1; a.temp = 1.0f;
    a.h0 = state_in; a.first_tok = first_word; a.state_out = state_out; a.attn_w_out = attn_w_out;
    AC_TIMED("bah_greedy", st);
    rc = bah_launch(bah_greedy_kernel, B, bah_smem_bytes(1), st, a); if (rc) return rc;
    AC_LAUNCHED("bah_greedy_kernel");
    return AC_OK;
}

int ac_bah_beam(const ac_bah_t* d, const float* fc_emb, const float* attn_emb, const int64_t* lens, const int64_t* tags,
                int B, int T, int max_len, int beam, float temp, int start_idx, int end_idx, int64_t* seq,
                void* ws, size_t ws_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 0 && T >= 1 && T <= kBahMaxT && max_len >= 1 && max_len <= kBahMaxLen,
               "ac_bah_beam: sizes out of range (B=%d T=%d max_len=%d)", B, T, max_len);
    AC_REQUIRE(beam >= 1 && beam <= 5, "ac_bah_beam: beam size %d not in 1..5", beam);
    AC_REQUIRE(temp > 0.0f, "ac_bah_beam: temperature must be positive");
    if (B == 0) return AC_OK;
    AC_REQUIRE(d && fc_emb && attn_emb && lens && tags && seq, "ac_bah_beam: null argument");
    AC_REQUIRE(ws && ws_bytes >= ac_bah_workspace_bytes(d, B * beam, T), "ac_bah_beam: workspace too small (%zu < %zu)", ws_bytes,
               ac_bah_workspace_bytes(d, B * beam, T));
    cudaStream_t st = (cudaStream_t)stream;
    BahArgs a{};
    int rc = bah_prepare(d, fc_emb, attn_emb, B, T, (float*)ws, B * beam, a, st); if (rc) return rc;
    a.mem_len = lens; a.tags = tags; a.max_len = max_len; a.start_idx = start_idx; a.end_idx = end_idx;
    a.seq = seq; a.logprob = nullptr; a.logit_out = nullptr; a.beam = beam; a.temp = temp;
    AC_TIMED("bah_beam", st);
#define AC_BAH_CASE(RR) case RR: rc = bah_launch(bah_beam_kernel<RR>, B, bah_smem_bytes(RR), st, a); break;
    switch (beam) { AC_BAH_CASE(1) AC_BAH_CASE(2) AC_BAH_CASE(3) AC_BAH_CASE(4) AC_BAH_CASE(5) }
#undef AC_BAH_CASE
    if (rc) return rc;
    AC_LAUNCHED("bah_beam_kernel");
    return AC_OK;
}

}  // extern "C"
