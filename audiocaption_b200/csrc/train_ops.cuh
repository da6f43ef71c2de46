// Building blocks of the training step: Linear forward / backward on the tcgen05 GEMM,
// LayerNorm, multi-head attention, embedding, dropout -- forward AND backward, fp32 (3xTF32 GEMMs), no autograd.
// Used by trm_train.cu (Transformer decoder) and bigru_train.cu (bi-GRU encoder); the loss and the optimizer live in
// train_ops.cu as well.
#pragma once
#include <vector>

#include "gemm.cuh"

namespace ac {

// ------------------------------------------------------------------------------------ counter-based dropout RNG
// keep(seed, site, index) is a pure function, so the backward pass regenerates the mask instead of storing it.
#ifdef __CUDACC__
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// multiplier of an element under dropout(p): 0 with probability p, else 1 / (1 - p);  p == 0 -> 1
__device__ __forceinline__ float drop_scale(uint64_t seed, uint32_t site, uint64_t idx, float p) {
    if (p <= 0.0f) return 1.0f;
    const uint64_t h = mix64(mix64(seed ^ ((uint64_t)site << 40)) + idx);
    const float u = (float)(h >> 40) * (1.0f / 16777216.0f);      // 24 random bits -> [0, 1)
    return u < p ? 0.0f : 1.0f / (1.0f - p);
}
#endif

struct Dropout { float p = 0.0f; uint64_t seed = 0; };

// ------------------------------------------------------------------------------------ Linear layers
// y = act(x W^T + b): W [N, K] is a LIVE parameter (it changes every optimizer step); `refresh` re-packs it (and its
// transpose, for the input gradient) into the tensor core's layout -- once per step, before the forward pass.
struct Linear {
    const float* W = nullptr; const float* b = nullptr;   // parameters [N, K], [N] (b nullable)
    float* dW = nullptr; float* db = nullptr;             // gradients (nullable: frozen)
    int N = 0, K = 0;
    float* pk = nullptr; float* pkT = nullptr;            // packed W (forward) and W^T (input gradient), library-owned
    TcWeight tw, twT;
};
size_t linear_pack_floats(int N, int K, bool need_dx);     // storage `Linear::pk` (+ pkT) needs
int linear_refresh(Linear& l, bool need_dx, cudaStream_t st);
// Batched variant: plan every layer once (handle creation), upload the job table, then ONE tc_pack_multi launch per step.
long long linear_plan(Linear& l, bool need_dx, long long first, std::vector<TcPackJob>& jobs);
// Y [M, ldy] = act(X [M, K] W^T + b) (+ R);  act as in gemm.cuh
int linear_fwd(const Linear& l, const float* X, int M, float* Y, int ldy, int act, const float* R, cudaStream_t st);
// Gradients of y = x W^T + b given dY [M, N] (row stride ldy):
//   db = colsum(dY), dW = dY^T X (written, not accumulated), dX = dY W + R (R nullable; dX nullable).
// scratch: linear_bwd_scratch_floats(M, N, K) floats, 128-byte aligned.
size_t linear_bwd_scratch_floats(int M, int N, int K);
// Weight gradients are LEAVES of the backward pass (only the optimizer reads them) and their GEMMs are tiny (4-24 CTAs, a
// K loop over the token rows), so they leave the critical path: with `side` non-null the main stream only transposes dY
// into the call's PRIVATE scratch, and db (row sums of dY^T), the pack of X and the dW GEMM run on one of the side
// streams, concurrently with the dX GEMM and whatever follows on the main stream.  X must stay unmodified until
// SideStreams::join (saved forward activations are).  Without `side` everything runs on `st` in order.
struct SideStreams {
    static constexpr int kN = 2;
    cudaStream_t s[kN] = {};
    cudaEvent_t fork_ev[kN] = {}, done_ev[kN] = {}, mark_ev = nullptr;
    bool pending[kN] = {};
    int next = 0, last = -1;
    bool ready = false;
    int init();                                          // creates the streams / events (AC_TRAIN_SIDE=0: stays disabled)
    void destroy();
    bool enabled() const { return ready; }
    int fork(cudaStream_t main_st, cudaStream_t* side);  // *side waits for everything enqueued on main_st so far
    cudaStream_t last_stream() const { return last >= 0 ? s[last] : nullptr; }
    int mark();                                          // remember "everything enqueued on the last forked stream"
    int wait_mark(cudaStream_t main_st);                 // ... and make main_st wait for it
    int join(cudaStream_t main_st);                      // main_st waits for every side stream with work in flight
};
int linear_bwd(const Linear& l, const float* X, int ldx, const float* dY, int ldy, int M, float* dX, const float* R,
               float* scratch, cudaStream_t st, SideStreams* side = nullptr);

// ------------------------------------------------------------------------------------ small kernels
// out[n] = sum_m X[m * ld + n]  (deterministic order)
int colsum(const float* X, int M, int N, int ld, float* out, cudaStream_t st);
// out[n] = sum_m XT[n * Mp + m]  (Mp % 4 == 0; one warp per row, fixed order)
int rowsum(const float* XT, int N, int Mp, float* out, cudaStream_t st);
// XT [N, Mp] = X[M, N]^T (row stride ld), zero-padded to Mp >= M columns
int transpose_pad(const float* X, int M, int N, int ld, float* XT, int Mp, cudaStream_t st);
// Row-wise kernels take BASE pointers and a row range [row0, row0 + n_rows): dropout masks are functions of the absolute
// element index, so a forward pass run in two row ranges and a backward pass run over all rows agree.
// X0[m, :] = drop_pe( drop_in(emb[word[m]]) * scale + pe[m % L] )       (transformer_decoder.py:88-91)
int embed_fwd(const float* emb, const float* pe, const int64_t* word, int row0, int n_rows, int L, int D, int V, float scale,
              Dropout dp, float* X0, cudaStream_t st);
// demb[word[m], :] += dX0[m, :] * scale * masks     (atomic adds; demb zeroed by the caller)
int embed_bwd(const float* dX0, const int64_t* word, int n_rows, int L, int D, int V, float scale, Dropout dp, float* demb,
              cudaStream_t st);
// S = X + drop(O) (X nullable);  Y = LayerNorm(S) * gamma + beta;  saves S (nullable == in place over O not allowed), mean, rstd
int add_ln_fwd(const float* X, const float* O, const float* gamma, const float* beta, int row0, int M, int D, Dropout dp,
               uint32_t site, float* S, float* mean, float* rstd, float* Y, cudaStream_t st);
// dS = LayerNorm backward of dY (+ dYres added to the result when non-null: the residual path of the NEXT op);
// dO = dS * dropmask (nullable -> not written); dgamma / dbeta written (deterministic two-stage reduction).
// scratch: ln_bwd_scratch_floats(M, D).
size_t ln_bwd_scratch_floats(int M, int D);
int add_ln_bwd(const float* dY, const float* S, const float* mean, const float* rstd, const float* gamma, int M, int D,
               Dropout dp, uint32_t site, float* dS, float* dO, float* dgamma, float* dbeta, float* scratch, cudaStream_t st);
// in place: X[i] *= dropout mask(site, i) for i in [i0, i0 + n)
int dropout_apply(float* X, int64_t i0, int64_t n, Dropout dp, uint32_t site, cudaStream_t st);
// dX = dY * (Y > 0) * dropout mask   (Y = the saved post-ReLU, post-dropout activation; in place over dY allowed)
int relu_drop_bwd(const float* dY, const float* Y, int64_t n, Dropout dp, uint32_t site, float* dX, cudaStream_t st);
// dst[i, :] = src[rows[i], :] ;  dst[rows[i], :] (+)= src[i, :]
int gather_rows(const float* src, const int* rows, int n, int D, float* dst, cudaStream_t st);
int scatter_rows(const float* src, const int* rows, int n, int D, float* dst, cudaStream_t st);
// idx[m] = argmax_v X[m * ld + v] (first maximum), logprob[m] = max log-softmax (nullable)
int argmax_rows(const float* X, int ld, int M, int V, int64_t* idx, float* logprob, cudaStream_t st);

// Multi-head attention core on projected tensors (nn.MultiheadAttention, batch-major rows m = seq * L + t).
//   Q rows [n_seq * L] (row stride ldq), K / V rows [n_kv_seq * Lk] (row stride ldkv); head h = columns [h*64, h*64+64).
//   kv sequence of query sequence s is s % n_kv_seq (the memory of a clip is shared by its GT and sampled rows).
//   mask: causal (j <= i) when `causal`; key j of sequence s masked when key_pad[s * Lk + j] != 0 (nullable) or
//   j >= kv_len[s % n_kv_seq] (nullable).
//   P [n_seq, H, L, Lk] = softmax(QK^T / 8 + mask) is saved; O = drop(P) V.
struct AttnArgs {
    const float* Q; const float* K; const float* V; int ldq, ldkv;
    const unsigned char* key_pad; const int64_t* kv_len;
    int seq0, n_seq, n_kv_seq, L, Lk, H; bool causal;   // sequences [seq0, seq0 + n_seq) of base pointers Q / P / O / key_pad
    Dropout dp; uint32_t site;
    float* P; float* O; int ldo;
};
int attn_fwd(const AttnArgs& a, cudaStream_t st);
// dQ / dK / dV from dO (same geometry).  dK / dV of a kv sequence accumulate over the query sequences that share it.
struct AttnBwdArgs {
    AttnArgs f;            // forward geometry + saved P (f.O unused)
    const float* dO; int lddo;
    float* dQ; int lddq; float* dK; float* dV; int lddkv;
};
int attn_bwd(const AttnBwdArgs& a, cudaStream_t st);

constexpr int kAttnHeadDim = 64;
constexpr int kAttnMaxL = 64;       // query positions per sequence, forward (PE table: 100; captions are <= 22 tokens, text_tokenizer.py:46-47)
constexpr int kAttnMaxLBwd = 32;    // query positions per sequence, backward (training: cap length - 1 <= 21)
constexpr int kAttnMaxLk = 128;     // keys per sequence

}  // namespace ac
