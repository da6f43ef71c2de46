// Transformer caption decoder, dense full-prefix forward AND backward (sm_100a).
//
// Replaces captioning/models/transformer_decoder.py:80-103 `TransformerDecoder.forward` (HF copy hf_wrapper.py:1045-1068)
// as called by the training loop: `TransformerModel.seq_forward` (transformer_model.py:20-32, teacher forcing) and the
// scheduled-sampling `stepwise_forward(mode="train")` (captioning/models/base.py:131-170, transformer_model.py:34-57).
//
//   memory:  P = LayerNorm(Dropout(ReLU(attn_emb W0^T + b0)))                                   (attn_proj)
//   tokens:  X0 = Dropout(Dropout(emb[word]) * sqrt(d) + PE)
//   per layer (nn.TransformerDecoderLayer, post-norm, ReLU):
//            X1 = LN1(X  + Dropout(SelfAttn(X;  causal + key padding)))
//            X2 = LN2(X1 + Dropout(CrossAttn(X1, P; memory key padding)))
//            X3 = LN3(X2 + Dropout(W2 Dropout(ReLU(W1 X2 + b1)) + b2))
//   logits = X W_cls^T
//
// Because the self-attention is causal, position t of one full-prefix pass over a token row equals the LAST position of
// the reference's step-t call on the prefix [:t+1]; the scheduled-sampling loop therefore needs at most two token rows
// per clip -- the ground-truth caption and the model's own samples -- and the step-t logits are taken from one or the
// other according to that step's coin (host side: audiocaption_b200/captioning/models/transformer_model.py).  Sequences
// [0, n_seq) share the memory of clip (seq % B).
//
// Every Linear runs on the tcgen05 GEMM (gemm_tc.cu, 3xTF32 = fp32-level accuracy) in all three roles (y = xW^T,
// dx = dy W, dW = dy^T x); weights are re-packed from the live parameters once per step (ac_trm_train_refresh).
// Activations needed by the backward pass live in the caller's workspace; dropout masks are regenerated from a
// counter-based RNG.  All row-major fp32, rows m = seq * L + t.
#include <algorithm>
#include <vector>

#include "train_ops.cuh"

namespace ac {

struct TrmTrainLayer {
    Linear sa_in, sa_out, ca_q, ca_kv, ca_out, l1, l2;
    const float* nw[3]; const float* nb[3];
    float* dnw[3]; float* dnb[3];
};

// dropout sites (distinct RNG streams)
enum { SITE_EMB_IN = 0, SITE_EMB_PE = 1, SITE_MEM = 2, SITE_LAYER0 = 8 };
enum { LS_SA_P = 0, LS_SA_OUT = 1, LS_CA_P = 2, LS_CA_OUT = 3, LS_FF_H = 4, LS_FF_OUT = 5, LS_STRIDE = 8 };

__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int rows_p, int D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows_p * D) return;
    dst[i] = i < (int64_t)rows * D ? src[i] : 0.0f;
}
__global__ void add_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

}  // namespace ac

struct ac_trm_train {
    int D = 0, H = 0, NL = 0, FF = 0, V = 0, Vp = 0, E = 0, pe_len = 0;
    bool tied = false;
    const float* emb = nullptr; float* demb = nullptr; const float* pe = nullptr;
    const float* cls_w = nullptr; float* dcls_w = nullptr;
    std::vector<ac::TrmTrainLayer> layer;
    ac::Linear cls, ap0;
    const float* ap_lnw = nullptr; const float* ap_lnb = nullptr; float* dap_lnw = nullptr; float* dap_lnb = nullptr;
    float* cls_stage = nullptr; float* dcls_stage = nullptr;     // [Vp, D] when V % 8 != 0
    float* blob = nullptr;
    ac::TcPackJob* jobs_dev = nullptr; int n_jobs = 0; long long job_items = 0;   // one batched re-pack launch per step
    ac::SideStreams side;                                          // weight-gradient GEMMs off the backward pass's critical path
};

namespace ac {

// Workspace layout (floats).  n_seq = maximum number of token rows (2 x clips with scheduled sampling).
struct TrmWs {
    size_t total = 0;
    // memory side
    size_t U, Pm, m_mean, m_rstd; std::vector<size_t> KV;
    // token side
    size_t X0;
    struct L { size_t QKV, Psa, Asa, S1, mean1, rstd1, X1, Qc, Pca, Aca, S2, mean2, rstd2, X2, Hff, S3, mean3, rstd3, X3; };
    std::vector<L> layer;
    size_t Xsel;
    // backward scratch
    size_t dA, dB, dC, dQKV, dH, dKV, dPm, dU, ln;
    size_t side, side_floats;    // private scratch of every linear_bwd call of one backward pass (they overlap in time)
};
static TrmWs trm_ws_layout(const ac_trm_train* h, int n_seq, int L, int B, int T) {
    TrmWs w;
    const size_t M = (size_t)n_seq * L, Mm = (size_t)B * T;
    const int D = h->D;
    auto take = [&](size_t n) { size_t o = w.total; w.total += align_up(n, 32); return o; };
    w.U = take(Mm * D); w.Pm = take(Mm * D); w.m_mean = take(Mm); w.m_rstd = take(Mm);
    for (int l = 0; l < h->NL; ++l) w.KV.push_back(take(Mm * 2 * D));
    w.X0 = take(M * D);
    for (int l = 0; l < h->NL; ++l) {
        TrmWs::L a;
        a.QKV = take(M * 3 * D); a.Psa = take((size_t)n_seq * h->H * L * L); a.Asa = take(M * D); a.S1 = take(M * D);
        a.mean1 = take(M); a.rstd1 = take(M); a.X1 = take(M * D);
        a.Qc = take(M * D); a.Pca = take((size_t)n_seq * h->H * L * T); a.Aca = take(M * D); a.S2 = take(M * D);
        a.mean2 = take(M); a.rstd2 = take(M); a.X2 = take(M * D);
        a.Hff = take(M * h->FF); a.S3 = take(M * D); a.mean3 = take(M); a.rstd3 = take(M); a.X3 = take(M * D);
        w.layer.push_back(a);
    }
    w.Xsel = take(M * D);
    w.dA = take(M * D); w.dB = take(M * D); w.dC = take(M * D); w.dQKV = take(M * 3 * D); w.dH = take(M * h->FF);
    w.dKV = take(Mm * 2 * D); w.dPm = take(Mm * D); w.dU = take(Mm * D);
    const size_t Mmax = std::max(M, Mm);
    w.ln = take(ln_bwd_scratch_floats((int)Mmax, D));
    // one slot per linear_bwd call, in the order ac_trm_train_bwd takes them
    size_t side = linear_bwd_scratch_floats((int)M, h->Vp, D) + linear_bwd_scratch_floats((int)Mm, D, h->E);
    for (int l = 0; l < h->NL; ++l)
        side += linear_bwd_scratch_floats((int)M, D, h->FF) + linear_bwd_scratch_floats((int)M, h->FF, D) +
                3 * linear_bwd_scratch_floats((int)M, D, D) + linear_bwd_scratch_floats((int)Mm, 2 * D, D) +
                linear_bwd_scratch_floats((int)M, 3 * D, D);
    w.side_floats = side;
    w.side = take(side);
    return w;
}

}  // namespace ac

extern "C" {

int ac_trm_train_num_tensors(int nlayers) { return 2 + 18 * nlayers + 5; }

// params_dev: the tensors of ac_trm_create's order (word_embedding.weight, pos_encoder.pe, per layer 18, classifier.weight,
// attn_proj.0.weight, attn_proj.0.bias, attn_proj.3.weight, attn_proj.3.bias) -- LIVE parameter storage, read at every
// ac_trm_train_refresh / forward; grads_dev: matching gradient buffers (entry NULL = frozen; pos_encoder.pe is always frozen).
int ac_trm_train_create(const float* const* p, float* const* g, const int64_t* numels, int n_tensors, int d_model, int nhead,
                        int nlayers, int dim_ff, int vocab, int attn_emb_dim, int pe_len, void* stream, ac_trm_train_t** out) {
    using namespace ac;
    (void)stream;
    AC_REQUIRE(p && g && numels && out, "ac_trm_train_create: null argument");
    AC_REQUIRE(n_tensors == ac_trm_train_num_tensors(nlayers), "ac_trm_train_create: expected %d tensors, got %d",
               ac_trm_train_num_tensors(nlayers), n_tensors);
    AC_REQUIRE(d_model == 256 && nhead * kAttnHeadDim == d_model, "ac_trm_train_create: d_model %d / nhead %d not supported "
               "(the attention kernels are built for 64-wide heads, LayerNorm for width 256)", d_model, nhead);
    AC_REQUIRE(dim_ff % 8 == 0 && attn_emb_dim % 8 == 0 && vocab >= 8, "ac_trm_train_create: bad sizes");
    const int D = d_model, FF = dim_ff, V = vocab, E = attn_emb_dim;
    AC_REQUIRE(numels[0] == (int64_t)V * D && numels[1] == (int64_t)pe_len * D, "ac_trm_train_create: embedding / pe size mismatch");
    ac_trm_train_t* h = new ac_trm_train_t();
    h->D = D; h->H = nhead; h->NL = nlayers; h->FF = FF; h->V = V; h->Vp = (V + 7) / 8 * 8; h->E = E; h->pe_len = pe_len;
    h->emb = p[0]; h->demb = g[0]; h->pe = p[1];
    int i = 2;
    size_t pk_total = 0;
    auto lin = [&](Linear& L, const float* W, const float* b, float* dW, float* db, int N, int K, bool need_dx) {
        L.W = W; L.b = b; L.dW = dW; L.db = db; L.N = N; L.K = K;
        pk_total += linear_pack_floats(N, K, need_dx);
    };
    for (int l = 0; l < nlayers; ++l) {
        TrmTrainLayer y{};
        const int64_t want[18] = {(int64_t)3 * D * D, 3 * D, (int64_t)D * D, D, (int64_t)3 * D * D, 3 * D, (int64_t)D * D, D,
                                  (int64_t)FF * D, FF, (int64_t)D * FF, D, D, D, D, D, D, D};
        for (int k = 0; k < 18; ++k)
            if (numels[i + k] != want[k]) { delete h; set_error("ac_trm_train_create: layer %d tensor %d has %lld elements, expected %lld", l, k, (long long)numels[i + k], (long long)want[k]); return AC_ERR_ARG; }
        lin(y.sa_in, p[i], p[i + 1], g[i], g[i + 1], 3 * D, D, true);
        lin(y.sa_out, p[i + 2], p[i + 3], g[i + 2], g[i + 3], D, D, true);
        // multihead_attn.in_proj: rows [0, D) project the queries (tokens), rows [D, 3D) the keys / values (memory)
        lin(y.ca_q, p[i + 4], p[i + 5], g[i + 4], g[i + 5], D, D, true);
        lin(y.ca_kv, p[i + 4] + (size_t)D * D, p[i + 5] + D, g[i + 4] ? g[i + 4] + (size_t)D * D : nullptr,
            g[i + 5] ? g[i + 5] + D : nullptr, 2 * D, D, true);
        lin(y.ca_out, p[i + 6], p[i + 7], g[i + 6], g[i + 7], D, D, true);
        lin(y.l1, p[i + 8], p[i + 9], g[i + 8], g[i + 9], FF, D, true);
        lin(y.l2, p[i + 10], p[i + 11], g[i + 10], g[i + 11], D, FF, true);
        for (int k = 0; k < 3; ++k) { y.nw[k] = p[i + 12 + 2 * k]; y.nb[k] = p[i + 13 + 2 * k]; y.dnw[k] = g[i + 12 + 2 * k]; y.dnb[k] = g[i + 13 + 2 * k]; }
        h->layer.push_back(y);
        i += 18;
    }
    if (numels[i] != (int64_t)V * D || numels[i + 1] != (int64_t)D * E || numels[i + 2] != D || numels[i + 3] != D || numels[i + 4] != D) {
        delete h; set_error("ac_trm_train_create: classifier / attn_proj size mismatch"); return AC_ERR_ARG;
    }
    h->cls_w = p[i]; h->dcls_w = g[i];
    h->tied = p[i] == p[0];
    lin(h->ap0, p[i + 1], p[i + 2], g[i + 1], g[i + 2], D, E, false);      // the encoder gradient comes from a separate GEMM
    pk_total += align_up(tc_packed_floats(E, D), 32);                      // ... whose transposed pack lives here
    h->ap_lnw = p[i + 3]; h->ap_lnb = p[i + 4]; h->dap_lnw = g[i + 3]; h->dap_lnb = g[i + 4];
    const bool stage = h->Vp != V;
    lin(h->cls, h->cls_w, nullptr, h->dcls_w, nullptr, h->Vp, D, true);
    const size_t stage_floats = stage ? 2 * align_up((size_t)h->Vp * D, 32) : 0;
    int rc = check_cuda(cudaMalloc(&h->blob, (pk_total + stage_floats) * sizeof(float)), "ac_trm_train_create: cudaMalloc");
    if (rc != AC_OK) { delete h; return rc; }
    float* cur = h->blob;
    auto place = [&](Linear& L, bool need_dx) { L.pk = cur; cur += linear_pack_floats(L.N, L.K, need_dx); };
    for (auto& y : h->layer) { place(y.sa_in, true); place(y.sa_out, true); place(y.ca_q, true); place(y.ca_kv, true); place(y.ca_out, true); place(y.l1, true); place(y.l2, true); }
    place(h->ap0, false);
    h->ap0.pkT = cur; cur += align_up(tc_packed_floats(E, D), 32);
    place(h->cls, true);
    if (stage) {
        h->cls_stage = cur; cur += align_up((size_t)h->Vp * D, 32);
        h->dcls_stage = cur; cur += align_up((size_t)h->Vp * D, 32);
        h->cls.W = h->cls_stage; h->cls.dW = h->dcls_w ? h->dcls_stage : nullptr;
    }
    // the static job table of ac_trm_train_refresh: every weight image (forward and transposed) of every Linear
    std::vector<TcPackJob> jobs;
    long long first = 0;
    for (auto& y : h->layer) {
        Linear* ls[7] = {&y.sa_in, &y.sa_out, &y.ca_q, &y.ca_kv, &y.ca_out, &y.l1, &y.l2};
        for (Linear* l : ls) first = linear_plan(*l, true, first, jobs);
    }
    first = linear_plan(h->ap0, false, first, jobs);
    {
        TcPackJob j;
        first += tc_pack_plan(h->ap0.W, E, D, 1, E, h->ap0.pkT, first, &j, &h->ap0.twT);
        jobs.push_back(j);
    }
    first = linear_plan(h->cls, true, first, jobs);
    h->n_jobs = (int)jobs.size(); h->job_items = first;
    rc = check_cuda(cudaMalloc(&h->jobs_dev, jobs.size() * sizeof(TcPackJob)), "ac_trm_train_create: cudaMalloc jobs");
    if (rc == AC_OK) rc = check_cuda(cudaMemcpy(h->jobs_dev, jobs.data(), jobs.size() * sizeof(TcPackJob), cudaMemcpyHostToDevice), "jobs upload");
    if (rc != AC_OK) { cudaFree(h->blob); cudaFree(h->jobs_dev); delete h; return rc; }
    rc = h->side.init();
    if (rc != AC_OK) { ac_trm_train_destroy(h); return rc; }
    *out = h;
    return AC_OK;
}

void ac_trm_train_destroy(ac_trm_train_t* h) {
    if (!h) return;
    h->side.destroy();
    cudaFree(h->blob);
    cudaFree(h->jobs_dev);
    delete h;
}

size_t ac_trm_train_workspace_bytes(const ac_trm_train_t* h, int n_seq, int L, int B, int T) {
    if (!h) return 0;
    return ac::trm_ws_layout(h, n_seq, L, B, T).total * sizeof(float);
}

// Re-pack every weight from the live parameters (call once after each optimizer step, before the forward pass).
int ac_trm_train_refresh(ac_trm_train_t* h, void* stream) {
    using namespace ac;
    AC_REQUIRE(h, "ac_trm_train_refresh: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cls_stage != nullptr) {
        const int64_t n = (int64_t)h->Vp * h->D;
        pad_rows_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(h->cls_w, h->cls_stage, h->V, h->Vp, h->D);
        AC_LAUNCHED("pad_rows_kernel");
    }
    return tc_pack_multi(h->jobs_dev, h->n_jobs, h->job_items, st);       // all ~32 weight images in one launch
}

// Memory side: attn_emb_dev [B, T, E] -> projected memory and every layer's cross-attention keys / values (kept in the
// workspace for ac_trm_train_seq_fwd / ac_trm_train_bwd).
int ac_trm_train_memory_fwd(ac_trm_train_t* h, const float* attn_emb_dev, int B, int T, int n_seq_max, int L, float p_drop,
                            uint64_t seed, void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && attn_emb_dev && workspace_dev, "ac_trm_train_memory_fwd: null argument");
    const TrmWs w = trm_ws_layout(h, n_seq_max, L, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_trm_train_memory_fwd: workspace too small (%zu < %zu)", workspace_bytes, w.total * sizeof(float));
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const int Mm = B * T, D = h->D;
    const Dropout dp{p_drop, seed};
    AC_TIMED("span_trm_train_memory_fwd", st);      // wrapper span: contains the per-kernel timers below
    int rc = linear_fwd(h->ap0, attn_emb_dev, Mm, ws + w.U, D, ACT_RELU, nullptr, st); if (rc) return rc;
    rc = dropout_apply(ws + w.U, 0, (int64_t)Mm * D, dp, SITE_MEM, st); if (rc) return rc;
    rc = add_ln_fwd(nullptr, ws + w.U, h->ap_lnw, h->ap_lnb, 0, Mm, D, Dropout{}, 0, nullptr, ws + w.m_mean, ws + w.m_rstd, ws + w.Pm, st);
    if (rc) return rc;
    for (int l = 0; l < h->NL; ++l) {
        rc = linear_fwd(h->layer[l].ca_kv, ws + w.Pm, Mm, ws + w.KV[l], 2 * D, ACT_NONE, nullptr, st); if (rc) return rc;
    }
    return AC_OK;
}

// Token side: sequences [seq0, seq0 + n_seq) of word_dev [n_seq_max, L] (int64) / key_pad_dev [n_seq_max, L] (uint8, 1 =
// padding key: `cap_padding_mask`) through the decoder layers.  The final hidden states stay in the workspace
// (ac_trm_train_logits reads them).  attn_len_dev [B] int64 = valid memory frames per clip.
int ac_trm_train_seq_fwd(ac_trm_train_t* h, const int64_t* word_dev, const unsigned char* key_pad_dev, int seq0, int n_seq,
                         int n_seq_max, int L, const int64_t* attn_len_dev, int B, int T, float p_drop, uint64_t seed,
                         void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && word_dev && key_pad_dev && attn_len_dev && workspace_dev, "ac_trm_train_seq_fwd: null argument");
    AC_REQUIRE(seq0 >= 0 && n_seq >= 0 && seq0 + n_seq <= n_seq_max && L >= 1 && L <= h->pe_len, "ac_trm_train_seq_fwd: bad sequence range");
    AC_REQUIRE(n_seq_max % B == 0 && seq0 % B == 0 && n_seq % B == 0, "ac_trm_train_seq_fwd: token rows must come in multiples of the %d clips", B);
    const TrmWs w = trm_ws_layout(h, n_seq_max, L, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_trm_train_seq_fwd: workspace too small");
    if (n_seq == 0) return AC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const int D = h->D, FF = h->FF, r0 = seq0 * L, M = n_seq * L;
    const Dropout dp{p_drop, seed};
    AC_TIMED("span_trm_train_seq_fwd", st);      // wrapper span: contains the per-kernel timers below
    int rc = embed_fwd(h->emb, h->pe, word_dev, r0, M, L, D, h->V, sqrtf((float)D), dp, ws + w.X0, st); if (rc) return rc;
    const float* X = ws + w.X0;
    for (int l = 0; l < h->NL; ++l) {
        const TrmTrainLayer& y = h->layer[l];
        const TrmWs::L& a = w.layer[l];
        const uint32_t site = SITE_LAYER0 + l * LS_STRIDE;
        const size_t ro = (size_t)r0;
        // self-attention
        rc = linear_fwd(y.sa_in, X + ro * D, M, ws + a.QKV + ro * 3 * D, 3 * D, ACT_NONE, nullptr, st); if (rc) return rc;
        AttnArgs sa{};
        sa.Q = ws + a.QKV; sa.K = ws + a.QKV + D; sa.V = ws + a.QKV + 2 * D; sa.ldq = 3 * D; sa.ldkv = 3 * D;
        sa.key_pad = key_pad_dev; sa.kv_len = nullptr; sa.seq0 = seq0; sa.n_seq = n_seq; sa.n_kv_seq = n_seq_max; sa.L = L; sa.Lk = L;
        sa.H = h->H; sa.causal = true; sa.dp = dp; sa.site = site + LS_SA_P; sa.P = ws + a.Psa; sa.O = ws + a.Asa; sa.ldo = D;
        rc = attn_fwd(sa, st); if (rc) return rc;
        rc = linear_fwd(y.sa_out, ws + a.Asa + ro * D, M, ws + a.S1 + ro * D, D, ACT_NONE, nullptr, st); if (rc) return rc;   // S1 holds O for now
        rc = add_ln_fwd(X, ws + a.S1, y.nw[0], y.nb[0], r0, M, D, dp, site + LS_SA_OUT, ws + a.S1, ws + a.mean1, ws + a.rstd1, ws + a.X1, st);
        if (rc) return rc;
        // cross-attention over the clip's memory
        rc = linear_fwd(y.ca_q, ws + a.X1 + ro * D, M, ws + a.Qc + ro * D, D, ACT_NONE, nullptr, st); if (rc) return rc;
        AttnArgs ca{};
        ca.Q = ws + a.Qc; ca.K = ws + w.KV[l]; ca.V = ws + w.KV[l] + D; ca.ldq = D; ca.ldkv = 2 * D;
        ca.key_pad = nullptr; ca.kv_len = attn_len_dev; ca.seq0 = seq0; ca.n_seq = n_seq; ca.n_kv_seq = B; ca.L = L; ca.Lk = T;
        ca.H = h->H; ca.causal = false; ca.dp = dp; ca.site = site + LS_CA_P; ca.P = ws + a.Pca; ca.O = ws + a.Aca; ca.ldo = D;
        rc = attn_fwd(ca, st); if (rc) return rc;
        rc = linear_fwd(y.ca_out, ws + a.Aca + ro * D, M, ws + a.S2 + ro * D, D, ACT_NONE, nullptr, st); if (rc) return rc;
        rc = add_ln_fwd(ws + a.X1, ws + a.S2, y.nw[1], y.nb[1], r0, M, D, dp, site + LS_CA_OUT, ws + a.S2, ws + a.mean2, ws + a.rstd2, ws + a.X2, st);
        if (rc) return rc;
        // feed-forward
        rc = linear_fwd(y.l1, ws + a.X2 + ro * D, M, ws + a.Hff + ro * FF, FF, ACT_RELU, nullptr, st); if (rc) return rc;
        rc = dropout_apply(ws + a.Hff, (int64_t)ro * FF, (int64_t)M * FF, dp, site + LS_FF_H, st); if (rc) return rc;
        rc = linear_fwd(y.l2, ws + a.Hff + ro * FF, M, ws + a.S3 + ro * D, D, ACT_NONE, nullptr, st); if (rc) return rc;
        rc = add_ln_fwd(ws + a.X2, ws + a.S3, y.nw[2], y.nb[2], r0, M, D, dp, site + LS_FF_OUT, ws + a.S3, ws + a.mean3, ws + a.rstd3, ws + a.X3, st);
        if (rc) return rc;
        X = ws + a.X3;
    }
    return AC_OK;
}

// Classifier over selected hidden rows: logits_dev [n_rows, ac_trm_train_vocab_padded] = X[rows] W_cls^T (columns >= vocab
// are zero); rows_dev [n_rows] int32 indexes the hidden states of the last ac_trm_train_seq_fwd calls (NULL = rows
// 0..n_rows-1); embed_dev (nullable) [n_rows, d_model] receives X[rows] (`embed` of the reference's output dict).
// keep != 0 stores X[rows] in the workspace for ac_trm_train_bwd.
int ac_trm_train_logits(ac_trm_train_t* h, const int* rows_dev, int n_rows, int n_seq_max, int L, int B, int T, float* logits_dev,
                        float* embed_dev, int keep, void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && logits_dev && workspace_dev && n_rows >= 0 && n_rows <= n_seq_max * L, "ac_trm_train_logits: bad argument");
    const TrmWs w = trm_ws_layout(h, n_seq_max, L, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_trm_train_logits: workspace too small");
    if (n_rows == 0) return AC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const float* Xf = ws + w.layer[h->NL - 1].X3;
    const float* X = Xf;
    AC_TIMED("span_trm_train_logits", st);      // wrapper span: contains the per-kernel timers below
    int rc = AC_OK;
    if (rows_dev != nullptr || keep) {
        float* dst = keep ? ws + w.Xsel : ws + w.dA;
        if (rows_dev != nullptr) rc = gather_rows(Xf, rows_dev, n_rows, h->D, dst, st);
        else rc = check_cuda(cudaMemcpyAsync(dst, Xf, (size_t)n_rows * h->D * sizeof(float), cudaMemcpyDeviceToDevice, st), "ac_trm_train_logits copy");
        if (rc) return rc;
        X = dst;
    }
    if (embed_dev != nullptr)
        AC_CUDA(cudaMemcpyAsync(embed_dev, X, (size_t)n_rows * h->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return linear_fwd(h->cls, X, n_rows, logits_dev, h->Vp, ACT_NONE, nullptr, st);
}

int ac_trm_train_vocab_padded(const ac_trm_train_t* h) { return h ? h->Vp : 0; }

// Backward of everything above.  dlogits_dev [n_rows, Vp] (columns >= vocab must be zero) is the gradient of the logits
// ac_trm_train_logits(keep = 1) produced for rows_dev; n_seq token rows were run forward.  Writes the gradient of every
// parameter whose grads_dev entry is non-NULL (overwriting, not accumulating; the embedding gradient accumulates on top of
// the classifier's when the weights are tied) and, when dattn_emb_dev != NULL, the gradient of attn_emb [B, T, E].
int ac_trm_train_bwd(ac_trm_train_t* h, const float* dlogits_dev, const int* rows_dev, int n_rows, const int64_t* word_dev,
                     const unsigned char* key_pad_dev, int n_seq, int n_seq_max, int L, const float* attn_emb_dev,
                     const int64_t* attn_len_dev, int B, int T, float p_drop, uint64_t seed, float* dattn_emb_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && dlogits_dev && word_dev && key_pad_dev && attn_emb_dev && attn_len_dev && workspace_dev, "ac_trm_train_bwd: null argument");
    AC_REQUIRE(n_seq >= B && n_seq % B == 0 && n_seq <= n_seq_max, "ac_trm_train_bwd: bad sequence count");
    const TrmWs w = trm_ws_layout(h, n_seq_max, L, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_trm_train_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const int D = h->D, FF = h->FF, M = n_seq * L, Mm = B * T;
    const Dropout dp{p_drop, seed};
    float* lns = ws + w.ln;
    AC_TIMED("span_trm_train_bwd", st);      // wrapper span: contains the per-kernel timers below
    int rc = AC_OK;
    SideStreams* side = &h->side;
    size_t side_used = 0;                    // every linear_bwd call gets its own scratch: its dW work runs on a side stream
    auto slot = [&](int m, const Linear& y) {
        float* p = ws + w.side + side_used;
        side_used += linear_bwd_scratch_floats(m, y.N, y.K);
        return side_used <= w.side_floats ? p : nullptr;
    };
    float* lin = nullptr;
#define AC_SLOT(m, y) lin = slot(m, y); AC_REQUIRE(lin != nullptr, "ac_trm_train_bwd: side scratch exhausted")
    // ---- classifier: dXsel = dlogits W, dW = dlogits^T Xsel
    float* dXf = ws + w.dA;              // gradient of the current layer's output [M, D]
    if (rows_dev != nullptr) {
        AC_SLOT(n_rows, h->cls);
        rc = linear_bwd(h->cls, ws + w.Xsel, D, dlogits_dev, h->Vp, n_rows, ws + w.dB, nullptr, lin, st, side); if (rc) return rc;
        AC_CUDA(cudaMemsetAsync(dXf, 0, (size_t)M * D * sizeof(float), st));
        rc = scatter_rows(ws + w.dB, rows_dev, n_rows, D, dXf, st); if (rc) return rc;
    } else {
        AC_REQUIRE(n_rows == M, "ac_trm_train_bwd: without a row selection the logits must cover all %d rows", M);
        AC_SLOT(n_rows, h->cls);
        rc = linear_bwd(h->cls, ws + w.Xsel, D, dlogits_dev, h->Vp, n_rows, dXf, nullptr, lin, st, side); if (rc) return rc;
    }
    // the classifier's weight gradient was computed on a side stream (when enabled): the un-padding copy follows it there,
    // and the embedding gradient (which accumulates on top of it when the weights are tied) waits for the mark below
    const bool cls_on_side = side->enabled() && h->cls.dW != nullptr;
    if (h->dcls_stage != nullptr && h->dcls_w != nullptr)
        AC_CUDA(cudaMemcpyAsync(h->dcls_w, h->dcls_stage, (size_t)h->V * D * sizeof(float), cudaMemcpyDeviceToDevice,
                                cls_on_side ? side->last_stream() : st));
    if (cls_on_side) { rc = side->mark(); if (rc) return rc; }
    // ---- decoder layers, last to first.  dKV accumulates the memory-side gradient of each layer into dPm.
    bool first_mem = true;
    for (int l = h->NL - 1; l >= 0; --l) {
        const TrmTrainLayer& y = h->layer[l];
        const TrmWs::L& a = w.layer[l];
        const uint32_t site = SITE_LAYER0 + l * LS_STRIDE;
        const float* Xin = l == 0 ? ws + w.X0 : ws + w.layer[l - 1].X3;
        float* dS = ws + w.dB; float* dO = ws + w.dC;
        // X3 = LN3(X2 + drop(F)),  F = l2(Hff)
        rc = add_ln_bwd(dXf, ws + a.S3, ws + a.mean3, ws + a.rstd3, y.nw[2], M, D, dp, site + LS_FF_OUT, dS, dO, y.dnw[2], y.dnb[2], lns, st); if (rc) return rc;
        AC_SLOT(M, y.l2);
        rc = linear_bwd(y.l2, ws + a.Hff, FF, dO, D, M, ws + w.dH, nullptr, lin, st, side); if (rc) return rc;
        rc = relu_drop_bwd(ws + w.dH, ws + a.Hff, (int64_t)M * FF, dp, site + LS_FF_H, ws + w.dH, st); if (rc) return rc;
        AC_SLOT(M, y.l1);
        rc = linear_bwd(y.l1, ws + a.X2, D, ws + w.dH, FF, M, dXf, dS, lin, st, side); if (rc) return rc;           // dX2 = dH W1 + dS3
        // X2 = LN2(X1 + drop(Oc)),  Oc = ca_out(Aca)
        rc = add_ln_bwd(dXf, ws + a.S2, ws + a.mean2, ws + a.rstd2, y.nw[1], M, D, dp, site + LS_CA_OUT, dS, dO, y.dnw[1], y.dnb[1], lns, st); if (rc) return rc;
        AC_SLOT(M, y.ca_out);
        rc = linear_bwd(y.ca_out, ws + a.Aca, D, dO, D, M, dXf, nullptr, lin, st, side); if (rc) return rc;         // dXf = dAca
        AttnBwdArgs cb{};
        cb.f.Q = ws + a.Qc; cb.f.K = ws + w.KV[l]; cb.f.V = ws + w.KV[l] + D; cb.f.ldq = D; cb.f.ldkv = 2 * D;
        cb.f.key_pad = nullptr; cb.f.kv_len = attn_len_dev; cb.f.seq0 = 0; cb.f.n_seq = n_seq; cb.f.n_kv_seq = B; cb.f.L = L; cb.f.Lk = T;
        cb.f.H = h->H; cb.f.causal = false; cb.f.dp = dp; cb.f.site = site + LS_CA_P; cb.f.P = ws + a.Pca;
        cb.dO = dXf; cb.lddo = D; cb.dQ = dO; cb.lddq = D; cb.dK = ws + w.dKV; cb.dV = ws + w.dKV + D; cb.lddkv = 2 * D;
        rc = attn_bwd(cb, st); if (rc) return rc;                                                              // dO = dQc
        AC_SLOT(M, y.ca_q);
        rc = linear_bwd(y.ca_q, ws + a.X1, D, dO, D, M, dXf, dS, lin, st, side); if (rc) return rc;                  // dX1 = dQc Wq + dS2
        // memory side of this layer: dPm (+)= dKV W_kv
        AC_SLOT(Mm, y.ca_kv);
        if (first_mem) { rc = linear_bwd(y.ca_kv, ws + w.Pm, D, ws + w.dKV, 2 * D, Mm, ws + w.dPm, nullptr, lin, st, side); first_mem = false; }
        else {
            rc = linear_bwd(y.ca_kv, ws + w.Pm, D, ws + w.dKV, 2 * D, Mm, ws + w.dU, ws + w.dPm, lin, st, side);
            if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(ws + w.dPm, ws + w.dU, (size_t)Mm * D * sizeof(float), cudaMemcpyDeviceToDevice, st), "dPm copy");
        }
        if (rc) return rc;
        // X1 = LN1(X + drop(Osa)),  Osa = sa_out(Asa)
        rc = add_ln_bwd(dXf, ws + a.S1, ws + a.mean1, ws + a.rstd1, y.nw[0], M, D, dp, site + LS_SA_OUT, dS, dO, y.dnw[0], y.dnb[0], lns, st); if (rc) return rc;
        AC_SLOT(M, y.sa_out);
        rc = linear_bwd(y.sa_out, ws + a.Asa, D, dO, D, M, dXf, nullptr, lin, st, side); if (rc) return rc;         // dXf = dAsa
        AttnBwdArgs sb{};
        sb.f.Q = ws + a.QKV; sb.f.K = ws + a.QKV + D; sb.f.V = ws + a.QKV + 2 * D; sb.f.ldq = 3 * D; sb.f.ldkv = 3 * D;
        sb.f.key_pad = key_pad_dev; sb.f.kv_len = nullptr; sb.f.seq0 = 0; sb.f.n_seq = n_seq; sb.f.n_kv_seq = n_seq; sb.f.L = L; sb.f.Lk = L;
        sb.f.H = h->H; sb.f.causal = true; sb.f.dp = dp; sb.f.site = site + LS_SA_P; sb.f.P = ws + a.Psa;
        sb.dO = dXf; sb.lddo = D; sb.dQ = ws + w.dQKV; sb.lddq = 3 * D; sb.dK = ws + w.dQKV + D; sb.dV = ws + w.dQKV + 2 * D; sb.lddkv = 3 * D;
        rc = attn_bwd(sb, st); if (rc) return rc;
        AC_SLOT(M, y.sa_in);
        rc = linear_bwd(y.sa_in, Xin, D, ws + w.dQKV, 3 * D, M, dXf, dS, lin, st, side); if (rc) return rc;         // dXin = dQKV Win + dS1
    }
    // ---- embedding
    if (h->demb != nullptr) {
        if (cls_on_side && h->tied) { rc = side->wait_mark(st); if (rc) return rc; }
        if (!(h->tied && h->dcls_w != nullptr)) AC_CUDA(cudaMemsetAsync(h->demb, 0, (size_t)h->V * D * sizeof(float), st));
        rc = embed_bwd(dXf, word_dev, M, L, D, h->V, sqrtf((float)D), dp, h->demb, st); if (rc) return rc;
    }
    // ---- memory projection: Pm = LN(U), U = drop(relu(attn_emb W0^T + b0))
    rc = add_ln_bwd(ws + w.dPm, ws + w.U, ws + w.m_mean, ws + w.m_rstd, h->ap_lnw, Mm, D, Dropout{}, 0, ws + w.dU, nullptr, h->dap_lnw, h->dap_lnb, lns, st);
    if (rc) return rc;
    rc = relu_drop_bwd(ws + w.dU, ws + w.U, (int64_t)Mm * D, dp, SITE_MEM, ws + w.dU, st); if (rc) return rc;
    {
        Linear ap = h->ap0;
        if (dattn_emb_dev == nullptr) ap.pkT = nullptr;
        AC_SLOT(Mm, ap);
        rc = linear_bwd(ap, attn_emb_dev, h->E, ws + w.dU, D, Mm, dattn_emb_dev, nullptr, lin, st, side); if (rc) return rc;
    }
#undef AC_SLOT
    return side->join(st);          // the caller's stream sees every gradient (and may reuse the workspace) after this call
}

}  // extern "C"
