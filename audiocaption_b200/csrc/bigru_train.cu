// Bidirectional multi-layer GRU encoder: training forward (activations saved) and backward through time (sm_100a).
//
// Replaces, for the training step, captioning/models/rnn_encoder.py:34-49 `RnnEncoder.forward` (pack_wrapper(nn.GRU) with
// inter-layer dropout, captioning/utils/model_util.py:10-27) and its autograd backward (python_scripts/train_eval/
// run.py:124 `loss.backward()`).  Same packed-sequence semantics as bigru.cu.
//
// Forward per layer: the input projections of each direction as a tcgen05 GEMM from the LIVE weights (re-packed once per
// step), then bigru_recurrence_kernel<SAVE> (bigru.cu) which also records r, z, n, hn and the previous hidden state.
// Backward per layer, last to first:
//   1. bigru_bwd_kernel: the recurrence in reverse.  Same decomposition as the forward kernel -- one 8-CTA cluster per
//      (direction, 8 clips), CTA c owns hidden units [32c, 32c+32) and keeps its 96 rows of W_hh in shared memory -- but
//      the matvec is the transposed one: each CTA turns its 96 gate gradients into a partial dh over ALL 256 units and
//      the partials are reduce-scattered over distributed shared memory (CTA c' receives the 8 partials of its 32 units).
//      Emits dGi (gradient of W_ih x + b_ih) and dGh (gradient of W_hh h + b_hh) for every (clip, frame, direction).
//   2. GEMMs on the tensor cores: dW_ih = dGi^T X, dW_hh = dGh^T Hprev, dX = dGi W_ih (layers >= 1), bias gradients as
//      column sums; the inter-layer dropout mask is regenerated from the counter RNG.
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "bigru.cuh"
#include "train_ops.cuh"

namespace cg = cooperative_groups;

namespace ac {

constexpr uint32_t SITE_GRU_LAYER0 = 64;       // dropout stream of layer l's output: SITE_GRU_LAYER0 + l

struct GruBwdArgs {
    const float* dOut;      // [B, T, 2H] gradient of the layer output
    const float* save;      // [B, T, 2, 4, H]
    const float* hprev;     // [B, T, 2H]
    const float* whh[2];    // [3H, H]
    const int64_t* lens;
    float* dGi;             // [B*T, 6H]  (dir, gate r|z|n, unit)
    float* dGh;             // [B*T, 6H]
    int B, T;
};

constexpr size_t kGruBwdSmem = ((size_t)kGruH * kGruWtStride            // Wt [H k][97]: this CTA's 96 rows of W_hh, transposed
                                + kGruRows * kGruClips                  // dgh [96 j][8 clips]
                                + 2 * kGruCluster * kGruUnits * kGruClips   // recv [2][8 src][32 k][8 clips]
                                + kGruH * kGruClips) * sizeof(float);  // part [256 k][8 clips]

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
bigru_bwd_kernel(const GruBwdArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* Wt = gsm;                                           // Wt[k * 97 + j] = W_hh[row(j), k]
    float* dgh = Wt + kGruH * kGruWtStride;                    // [96][8]
    float* recv = dgh + kGruRows * kGruClips;                  // [2][8][32][8]
    float* part = recv + 2 * kGruCluster * kGruUnits * kGruClips;   // [256][8]
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int cl = blockIdx.x / kGruCluster;
    const int groups = (a.B + kGruClips - 1) / kGruClips;
    const int dir = cl / groups, b0 = (cl % groups) * kGruClips;
    const int tid = threadIdx.x;
    const float* W = a.whh[dir];
    for (int i = tid; i < kGruRows * kGruH; i += kGruThreads) {
        const int j = i / kGruH, k = i % kGruH;
        const int row = (j / kGruUnits) * kGruH + c * kGruUnits + (j % kGruUnits);
        Wt[k * kGruWtStride + j] = __ldg(W + (size_t)row * kGruH + k);
    }
    for (int i = tid; i < 2 * kGruCluster * kGruUnits * kGruClips; i += kGruThreads) recv[i] = 0.0f;
    const int u = tid % kGruUnits, bl = (tid / kGruUnits) % kGruClips;
    const bool cell = tid < kGruUnits * kGruClips;
    const int b = b0 + bl;
    const bool clip_ok = cell && b < a.B;
    const int unit = c * kGruUnits + u;
    const int len = clip_ok ? (int)min((int64_t)a.T, max((int64_t)0, a.lens[b])) : 0;
    float dh_carry = 0.0f;            // dh * z of the step processed before (the direct path h_{t-1} -> h_t)
    // matvec role (all 512 threads): (k, half of the clips)
    const int mk = tid % kGruH, mh = tid / kGruH;
    // publish role: float4 `pv` of the 1 KB slice [32 k][8 clips] destined to CTA `prk`
    const int prk = tid >> 6, pv = tid & 63;
    cluster.sync();

    for (int s = 0; s < a.T; ++s) {
        // forward order of direction 0 is t = 0..T-1, of direction 1 t = T-1..0; walk it backwards
        const int t = dir == 0 ? a.T - 1 - s : s;
        const float* rc = recv + (s & 1) * kGruCluster * kGruUnits * kGruClips;
        if (cell) {
            float gr = 0.0f, gz = 0.0f, gn = 0.0f, ghn = 0.0f;
            const bool active = t < len;
            if (active) {
                float dh = dh_carry;
#pragma unroll
                for (int src = 0; src < kGruCluster; ++src) dh += rc[(src * kGruUnits + u) * kGruClips + bl];   // W_hh^T dgh of the previous step
                const size_t o = ((size_t)b * a.T + t) * (2 * kGruH) + dir * kGruH + unit;
                dh += a.dOut[o];
                const float* sv = a.save + (((size_t)b * a.T + t) * 2 + dir) * 4 * kGruH + unit;
                const float r = sv[0], z = sv[kGruH], n = sv[2 * kGruH], hn = sv[3 * kGruH];
                const float hp = a.hprev[o];
                const float dn = dh * (1.0f - z);
                const float dz = dh * (hp - n);
                const float dan = dn * (1.0f - n * n);
                gr = dan * hn * r * (1.0f - r);
                gz = dz * z * (1.0f - z);
                gn = dan;
                ghn = dan * r;
                dh_carry = dh * z;
            } else {
                dh_carry = 0.0f;      // before the first / after the last valid frame nothing flows
            }
            if (clip_ok) {
                float* gi = a.dGi + ((size_t)b * a.T + t) * (6 * kGruH) + dir * 3 * kGruH + unit;
                float* gh = a.dGh + ((size_t)b * a.T + t) * (6 * kGruH) + dir * 3 * kGruH + unit;
                gi[0] = gr; gi[kGruH] = gz; gi[2 * kGruH] = gn;
                gh[0] = gr; gh[kGruH] = gz; gh[2 * kGruH] = ghn;
            }
            dgh[u * kGruClips + bl] = gr;
            dgh[(kGruUnits + u) * kGruClips + bl] = gz;
            dgh[(2 * kGruUnits + u) * kGruClips + bl] = ghn;
        }
        __syncthreads();
        {   // part[k][clips] = sum_j W_hh[row(j), k] * dgh[j][clips]   (this CTA's 96 rows)
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* wp = Wt + mk * kGruWtStride;
            const float4* gp = reinterpret_cast<const float4*>(dgh) + mh;
#pragma unroll 8
            for (int j = 0; j < kGruRows; ++j) {
                const float w = wp[j];
                const float4 g4 = gp[j * 2];
                acc[0] = fmaf(w, g4.x, acc[0]); acc[1] = fmaf(w, g4.y, acc[1]);
                acc[2] = fmaf(w, g4.z, acc[2]); acc[3] = fmaf(w, g4.w, acc[3]);
            }
            *reinterpret_cast<float4*>(part + mk * kGruClips + mh * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        {   // reduce-scatter: slice k in [32 prk, 32 prk + 32) goes to CTA prk, slot `c` of the other parity buffer
            float* dstbuf = recv + ((s + 1) & 1) * kGruCluster * kGruUnits * kGruClips + c * kGruUnits * kGruClips;
            const float4* src = reinterpret_cast<const float4*>(part + prk * kGruUnits * kGruClips);
            float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(dstbuf, prk));
            dst[pv] = src[pv];
        }
        cluster.sync();
    }
}

struct GruTrainLayer {
    Linear ih[2];                         // W_ih of each direction, [3H, Din]
    const float* whh[2]; const float* bhh[2];
    float* dwhh[2]; float* dbhh[2];
    int din;
};

__global__ void drop_mul_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, Dropout dp, uint32_t site) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * drop_scale(dp.seed, site, (uint64_t)i, dp.p);
}

}  // namespace ac

struct ac_bigru_train {
    int input_dim = 0, layers = 0;
    bool need_dx0 = false;            // also pack layer 0's W_ih^T (gradient w.r.t. the encoder input)
    ac::TcPackJob* jobs_dev = nullptr; int n_jobs = 0; long long job_items = 0;
    std::vector<ac::GruTrainLayer> layer;
    float* blob = nullptr;
    ac::SideStreams side;             // weight-gradient GEMMs off the backward pass's critical path
};

namespace ac {
struct GruWs {
    size_t total = 0;
    size_t G;                                     // [M, 6H] input projections of the current layer (forward only)
    std::vector<size_t> Y, Yd, save, hprev;       // per layer: output, dropped output (= next layer's input), saved gates
    size_t dGi, dGh, dY, dX, side, side_floats;   // side: private scratch of every linear_bwd call of one backward pass
};
static GruWs gru_ws_layout(const ac_bigru_train* h, int B, int T) {
    GruWs w;
    const size_t M = (size_t)B * T;
    const int H = kGruH;
    auto take = [&](size_t n) { size_t o = w.total; w.total += align_up(n, 32); return o; };
    w.G = take(M * 6 * H);
    for (int l = 0; l < h->layers; ++l) {
        w.Y.push_back(take(M * 2 * H)); w.Yd.push_back(take(M * 2 * H));
        w.save.push_back(take(M * 8 * H)); w.hprev.push_back(take(M * 2 * H));
    }
    w.dGi = take(M * 6 * H); w.dGh = take(M * 6 * H); w.dY = take(M * 2 * H); w.dX = take(M * 2 * H);
    size_t side = 0;
    for (int l = 0; l < h->layers; ++l)
        side += 2 * (linear_bwd_scratch_floats((int)M, 3 * H, l == 0 ? h->input_dim : 2 * H) + linear_bwd_scratch_floats((int)M, 3 * H, H));
    w.side_floats = side;
    w.side = take(side);
    return w;
}
}  // namespace ac

extern "C" {

// params_dev / grads_dev: nn.GRU order, 8 per layer (weight_ih, weight_hh, bias_ih, bias_hh, then `_reverse`); LIVE storage.
int ac_bigru_train_create(const float* const* p, float* const* g, const int64_t* numels, int n_tensors, int input_dim, int hidden,
                          int num_layers, int need_input_grad, void* stream, ac_bigru_train_t** out) {
    using namespace ac;
    (void)stream;
    AC_REQUIRE(p && g && numels && out, "ac_bigru_train_create: null argument");
    AC_REQUIRE(hidden == kGruH, "ac_bigru_train_create: hidden size %d is not supported (%d is)", hidden, kGruH);
    AC_REQUIRE(num_layers >= 1 && input_dim >= 8 && input_dim % 8 == 0, "ac_bigru_train_create: bad layer count / input size");
    AC_REQUIRE(n_tensors == num_layers * 8, "ac_bigru_train_create: expected %d tensors, got %d", num_layers * 8, n_tensors);
    const int H = kGruH;
    ac_bigru_train_t* h = new ac_bigru_train_t();
    h->input_dim = input_dim; h->layers = num_layers; h->need_dx0 = need_input_grad != 0;
    size_t pk_total = 0;
    for (int l = 0; l < num_layers; ++l) {
        const int din = l == 0 ? input_dim : 2 * H;
        GruTrainLayer L{};
        L.din = din;
        for (int d = 0; d < 2; ++d) {
            const int ti = (l * 2 + d) * 4;
            if (numels[ti] != (int64_t)3 * H * din || numels[ti + 1] != (int64_t)3 * H * H || numels[ti + 2] != 3 * H || numels[ti + 3] != 3 * H) {
                delete h; set_error("ac_bigru_train_create: layer %d direction %d has unexpected tensor sizes", l, d); return AC_ERR_ARG;
            }
            L.ih[d].W = p[ti]; L.ih[d].b = p[ti + 2]; L.ih[d].dW = g[ti]; L.ih[d].db = g[ti + 2]; L.ih[d].N = 3 * H; L.ih[d].K = din;
            L.whh[d] = p[ti + 1]; L.bhh[d] = p[ti + 3]; L.dwhh[d] = g[ti + 1]; L.dbhh[d] = g[ti + 3];
            pk_total += linear_pack_floats(3 * H, din, l > 0 || h->need_dx0);
        }
        h->layer.push_back(L);
    }
    int rc = check_cuda(cudaMalloc(&h->blob, pk_total * sizeof(float)), "ac_bigru_train_create: cudaMalloc");
    if (rc != AC_OK) { delete h; return rc; }
    float* cur = h->blob;
    for (int l = 0; l < num_layers; ++l)
        for (int d = 0; d < 2; ++d) { h->layer[l].ih[d].pk = cur; cur += linear_pack_floats(3 * H, h->layer[l].din, l > 0 || h->need_dx0); }
    std::vector<TcPackJob> jobs;
    long long first = 0;
    for (int l = 0; l < num_layers; ++l)
        for (int d = 0; d < 2; ++d) first = linear_plan(h->layer[l].ih[d], l > 0 || h->need_dx0, first, jobs);
    h->n_jobs = (int)jobs.size(); h->job_items = first;
    rc = check_cuda(cudaMalloc(&h->jobs_dev, jobs.size() * sizeof(TcPackJob)), "ac_bigru_train_create: cudaMalloc jobs");
    if (rc == AC_OK) rc = check_cuda(cudaMemcpy(h->jobs_dev, jobs.data(), jobs.size() * sizeof(TcPackJob), cudaMemcpyHostToDevice), "jobs upload");
    if (rc != AC_OK) { cudaFree(h->blob); cudaFree(h->jobs_dev); delete h; return rc; }
    rc = h->side.init();
    if (rc != AC_OK) { h->side.destroy(); cudaFree(h->blob); cudaFree(h->jobs_dev); delete h; return rc; }
    *out = h;
    return AC_OK;
}

void ac_bigru_train_destroy(ac_bigru_train_t* h) {
    if (!h) return;
    h->side.destroy();
    cudaFree(h->blob);
    cudaFree(h->jobs_dev);
    delete h;
}

size_t ac_bigru_train_workspace_bytes(const ac_bigru_train_t* h, int batch, int T) {
    if (!h) return 0;
    return ac::gru_ws_layout(h, batch, T).total * sizeof(float);
}

int ac_bigru_train_refresh(ac_bigru_train_t* h, void* stream) {
    using namespace ac;
    AC_REQUIRE(h, "ac_bigru_train_refresh: null handle");
    return tc_pack_multi(h->jobs_dev, h->n_jobs, h->job_items, (cudaStream_t)stream);
}

// x_dev [batch, T, input_dim] (T = max(lens): the frozen CNN's frames), lens_dev [batch] int64 -> out_dev [batch, T, 512].
// p_drop = nn.GRU's inter-layer dropout (applied to the output of every layer but the last), seed = this step's RNG seed.
int ac_bigru_train_fwd(ac_bigru_train_t* h, const float* x_dev, const int64_t* lens_dev, int B, int T, float p_drop, uint64_t seed,
                       float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && x_dev && lens_dev && out_dev && workspace_dev && B >= 1 && T >= 1, "ac_bigru_train_fwd: bad argument");
    const GruWs w = gru_ws_layout(h, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_bigru_train_fwd: workspace too small (%zu < %zu)", workspace_bytes, w.total * sizeof(float));
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const int H = kGruH, M = B * T;
    const Dropout dp{p_drop, seed};
    const float* in = x_dev;
    for (int l = 0; l < h->layers; ++l) {
        const GruTrainLayer& L = h->layer[l];
        // the two directions' input projections are independent: the second one runs on a side stream (forked BEFORE the
        // first is enqueued, so that it only waits for the layer's input)
        cudaStream_t ps = st;
        if (h->side.enabled()) { int rc = h->side.fork(st, &ps); if (rc) return rc; }
        for (int d = 0; d < 2; ++d) {
            int rc = linear_fwd(L.ih[d], in, M, ws + w.G + d * 3 * H, 6 * H, ACT_NONE, nullptr, d == 0 ? st : ps); if (rc) return rc;
        }
        { int rc = h->side.join(st); if (rc) return rc; }
        const bool last = l + 1 == h->layers;
        GruStepArgs a;
        a.G = ws + w.G; a.lens = lens_dev; a.out = last ? out_dev : ws + w.Y[l]; a.B = B; a.T_in = T; a.T_out = T;
        a.save = ws + w.save[l]; a.hprev = ws + w.hprev[l];
        for (int d = 0; d < 2; ++d) { a.whh[d] = L.whh[d]; a.bhh[d] = L.bhh[d]; }
        int rc = bigru_recurrence_launch(a, true, st); if (rc) return rc;
        if (!last) {
            if (p_drop > 0.0f) {
                const int64_t n = (int64_t)M * 2 * H;
                drop_mul_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(ws + w.Y[l], ws + w.Yd[l], n, dp, SITE_GRU_LAYER0 + l);
                AC_LAUNCHED("drop_mul_kernel");
                in = ws + w.Yd[l];
            } else {
                in = ws + w.Y[l];
            }
        }
    }
    return AC_OK;
}

// dout_dev [batch, T, 512] -> gradients of every GRU parameter (written into grads_dev) and, when dx_dev != NULL, of the input.
int ac_bigru_train_bwd(ac_bigru_train_t* h, const float* x_dev, const int64_t* lens_dev, const float* dout_dev, int B, int T,
                       float p_drop, uint64_t seed, float* dx_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(h && x_dev && lens_dev && dout_dev && workspace_dev && B >= 1 && T >= 1, "ac_bigru_train_bwd: bad argument");
    AC_REQUIRE(dx_dev == nullptr || h->need_dx0, "ac_bigru_train_bwd: the handle was created without need_input_grad");
    const GruWs w = gru_ws_layout(h, B, T);
    AC_REQUIRE(workspace_bytes >= w.total * sizeof(float), "ac_bigru_train_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const int H = kGruH, M = B * T;
    const Dropout dp{p_drop, seed};
    static cudaError_t attr_rc = cudaFuncSetAttribute(bigru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruBwdSmem);
    AC_CUDA(attr_rc);
    const float* dY = dout_dev;
    // weight-gradient GEMMs run on side streams (train_ops.cuh): every call has its own scratch; dGi / dGh are transposed on
    // the main stream before the next layer's recurrence overwrites them
    SideStreams* side = &h->side;
    size_t side_used = 0;
    auto slot = [&](const Linear& y) {
        float* p = ws + w.side + side_used;
        side_used += linear_bwd_scratch_floats(M, y.N, y.K);
        return side_used <= w.side_floats ? p : nullptr;
    };
    for (int l = h->layers - 1; l >= 0; --l) {
        const GruTrainLayer& L = h->layer[l];
        GruBwdArgs a;
        a.dOut = dY; a.save = ws + w.save[l]; a.hprev = ws + w.hprev[l]; a.lens = lens_dev; a.dGi = ws + w.dGi; a.dGh = ws + w.dGh;
        a.B = B; a.T = T;
        for (int d = 0; d < 2; ++d) a.whh[d] = L.whh[d];
        {
            const int groups = cdiv(B, kGruClips);
            AC_TIMED("bigru_bwd", st);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * groups * kGruCluster); cfg.blockDim = dim3(kGruThreads); cfg.dynamicSmemBytes = kGruBwdSmem; cfg.stream = st;
            AC_CUDA(cudaLaunchKernelEx(&cfg, bigru_bwd_kernel, a));
            AC_LAUNCHED("bigru_bwd_kernel");
        }
        const float* X = l == 0 ? x_dev : (p_drop > 0.0f ? ws + w.Yd[l - 1] : ws + w.Y[l - 1]);
        const int din = L.din;
        float* dXl = l == 0 ? dx_dev : ws + w.dX;
        for (int d = 0; d < 2; ++d) {
            // input side: db_ih, dW_ih (and dX for the first direction; the second adds onto it)
            const float* dGi = ws + w.dGi + d * 3 * H;
            int rc = AC_OK;
            // (no dense copy of this direction's gate gradients: every consumer reads them through the 6H row stride)
            float* lin = slot(L.ih[d]);
            AC_REQUIRE(lin != nullptr, "ac_bigru_train_bwd: side scratch exhausted");
            rc = linear_bwd(L.ih[d], X, din, dGi, 6 * H, M, nullptr, nullptr, lin, st, side); if (rc) return rc;
            // hidden side: db_hh, dW_hh = dGh^T Hprev
            const float* dGh = ws + w.dGh + d * 3 * H;
            if (L.dbhh[d] || L.dwhh[d]) {
                Linear hh; hh.N = 3 * H; hh.K = H; hh.dW = L.dwhh[d]; hh.db = L.dbhh[d];
                lin = slot(hh);
                AC_REQUIRE(lin != nullptr, "ac_bigru_train_bwd: side scratch exhausted");
                rc = linear_bwd(hh, ws + w.hprev[l] + d * H, 2 * H, dGh, 6 * H, M, nullptr, nullptr, lin, st, side); if (rc) return rc;
            }
        }
        if (dXl != nullptr) {
            // dX = dGi[:, dir 0] W_ih0 + dGi[:, dir 1] W_ih1: A operands read through the 6H row stride
            for (int d = 0; d < 2; ++d) {
                GemmArgs g;
                g.A = ws + w.dGi + d * 3 * H; g.C = d == 0 ? ws + w.dY : dXl; g.M = M; g.N = din; g.K = 3 * H;
                g.R = d == 0 ? nullptr : ws + w.dY; g.tw = &L.ih[d].twT; g.lda = 6 * H;
                AC_REQUIRE(L.ih[d].pkT != nullptr, "ac_bigru_train_bwd: layer %d has no transposed pack", l);
                int rc = gemm_tc(g, st); if (rc) return rc;
            }
            if (l > 0) {
                // the next (lower) layer's output gradient: through the inter-layer dropout mask
                const int64_t n = (int64_t)M * 2 * H;
                if (p_drop > 0.0f) {
                    drop_mul_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(dXl, ws + w.dY, n, dp, SITE_GRU_LAYER0 + (l - 1));
                    AC_LAUNCHED("drop_mul_kernel");
                    dY = ws + w.dY;
                } else {
                    dY = dXl;
                }
            }
        }
    }
    return side->join(st);
}

}  // extern "C"
