// Bidirectional multi-layer GRU encoder over padded, length-masked sequences (sm_100a).
//
// Replaces captioning/models/rnn_encoder.py:34-49 `RnnEncoder.forward` = `pack_wrapper(nn.GRU(batch_first, bidirectional,
// num_layers), x, lens)` (captioning/utils/model_util.py:10-27: sort by length, pack, run, pad with zeros, unsort), eval
// mode (no inter-layer dropout).  HF copy: hf_wrapper.py:1307-1347.
//
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)     z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn))  h' = (1 - z) * n + z * h
// Packed-sequence semantics: clip b runs over t < len[b] only (the reverse direction starts at t = len[b] - 1 from h = 0);
// outputs at t >= len[b] are zero.
//
// Per layer two launches:
//   1. the input projections of BOTH directions for all time steps as one tensor-core GEMM (gemm_tc, 3xTF32):
//      G[b*T + t, dir*3H + gate*H + u] = W_ih x + b_ih;
//   2. the recurrence: one 8-CTA cluster per (direction, group of 8 clips).  CTA c of a cluster owns hidden units
//      [32c, 32c+32): its 96 rows of W_hh stay in shared memory (97 KB, rows padded to 97 floats so that both the
//      transposing fill and the matvec reads are bank-conflict-free) for the whole sequence; the hidden state of the
//      group (8 x 256) is replicated in every CTA.  Per step: 384 threads compute the 96 x 8 dot products (row x clip
//      half x k half), 256 threads apply the gates, then the CTA's contiguous 1 KB state slice goes to the 7 peers as
//      128-bit distributed-shared-memory stores and one cluster barrier closes the step (4 us per step at 16 clips;
//      measured split before this layout: matvec 45 %, scalar remote stores 25 %, barrier 10 %).  Latency-bound by
//      construction (T sequential steps); the batch is spread over 2 x ceil(B/8) clusters.
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "bigru.cuh"

namespace cg = cooperative_groups;

namespace ac {

// SAVE (training, bigru_train.cu): also records what the backward pass needs -- the gates r, z, n, the hidden-side
// candidate pre-activation hn = W_hn h + b_hn, and the previous hidden state of every (clip, frame, direction).
template <bool SAVE>
__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
bigru_recurrence_kernel(const GruStepArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* Hs = gsm;                                  // [2][H k][8 clips]  double-buffered hidden state (16-byte aligned)
    float* pre = Hs + 2 * kGruH * kGruClips;          // [2 k-halves][96 j][8 clips]  partial W_hh h of this step
    float* Wt = pre + 2 * kGruRows * kGruClips;       // [H k][97]   (transposed slice of W_hh, padded rows)
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();          // owns hidden units [32c, 32c + 32)
    const int cl = blockIdx.x / kGruCluster;          // cluster index -> (direction, clip group)
    const int groups = (a.B + kGruClips - 1) / kGruClips;
    const int dir = cl / groups, b0 = (cl % groups) * kGruClips;
    const int tid = threadIdx.x;

    pdl_trigger();
    // constants first (weights), then wait for the projection GEMM
    const float* W = a.whh[dir];
    for (int i = tid; i < kGruRows * kGruH; i += kGruThreads) {
        const int j = i / kGruH, k = i % kGruH;       // coalesced along k; smem bank = (k + j) % 32: conflict-free
        const int row = (j / kGruUnits) * kGruH + c * kGruUnits + (j % kGruUnits);
        Wt[k * kGruWtStride + j] = __ldg(W + (size_t)row * kGruH + k);
    }
    for (int i = tid; i < 2 * kGruH * kGruClips; i += kGruThreads) Hs[i] = 0.0f;
    // cell-update role (threads 0..255): thread = (unit u, clip bl)
    const int u = tid % kGruUnits, bl = (tid / kGruUnits) % kGruClips;
    const bool cell = tid < kGruUnits * kGruClips;
    const int b = b0 + bl;
    const bool clip_ok = cell && b < a.B;
    const int unit = c * kGruUnits + u;
    const float bh_r = __ldg(a.bhh[dir] + unit), bh_z = __ldg(a.bhh[dir] + kGruH + unit), bh_n = __ldg(a.bhh[dir] + 2 * kGruH + unit);
    const int len = clip_ok ? (int)min((int64_t)a.T_out, max((int64_t)0, a.lens[b])) : 0;
    float h_own = 0.0f;
    // matvec role (threads 0..383): (row j, half of the clips, half of k)
    const int mj = tid % kGruRows, mq = tid / kGruRows;
    const int mh = mq & 1, kq = mq >> 1;
    // publish role (all 512 threads): float4 `pv` of this CTA's 1 KB state slice goes to peer `prk`
    const int prk = tid >> 6, pv = tid & 63;
    pdl_wait();
    cluster.sync();

    for (int s = 0; s < a.T_out; ++s) {
        const int t = dir == 0 ? s : a.T_out - 1 - s;
        const float* Hc = Hs + (s & 1) * kGruH * kGruClips;
        float* Hn = Hs + ((s + 1) & 1) * kGruH * kGruClips;
        // input-side pre-activations of this step: issue the loads before the matvec
        float gr = 0.f, gz = 0.f, gn = 0.f;
        const bool active = cell && t < len;
        if (active) {
            const float* g = a.G + ((size_t)b * a.T_in + t) * (6 * kGruH) + dir * 3 * kGruH + unit;
            gr = __ldg(g); gz = __ldg(g + kGruH); gn = __ldg(g + 2 * kGruH);
        }
        if (mq < 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* wp = Wt + (kq * (kGruH / 2)) * kGruWtStride + mj;
            const float4* hp = reinterpret_cast<const float4*>(Hc) + (kq * (kGruH / 2)) * 2 + mh;
#pragma unroll 8
            for (int k = 0; k < kGruH / 2; ++k) {
                const float w = wp[k * kGruWtStride];
                const float4 h4 = hp[k * 2];
                acc[0] = fmaf(w, h4.x, acc[0]); acc[1] = fmaf(w, h4.y, acc[1]);
                acc[2] = fmaf(w, h4.z, acc[2]); acc[3] = fmaf(w, h4.w, acc[3]);
            }
            *reinterpret_cast<float4*>(pre + (kq * kGruRows + mj) * kGruClips + mh * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        if (cell) {
            float h_new = h_own;
            if (active) {
                const float* p1 = pre + kGruRows * kGruClips;      // second k-half
                const float pr = pre[u * kGruClips + bl] + p1[u * kGruClips + bl];
                const float pz = pre[(kGruUnits + u) * kGruClips + bl] + p1[(kGruUnits + u) * kGruClips + bl];
                const float pn = pre[(2 * kGruUnits + u) * kGruClips + bl] + p1[(2 * kGruUnits + u) * kGruClips + bl];
                const float r = 1.0f / (1.0f + expf(-(gr + pr + bh_r)));
                const float z = 1.0f / (1.0f + expf(-(gz + pz + bh_z)));
                const float hn = pn + bh_n;
                const float n = tanhf(gn + r * hn);
                h_new = (1.0f - z) * n + z * h_own;
                if (SAVE) {
                    float* sv = a.save + (((size_t)b * a.T_out + t) * 2 + dir) * 4 * kGruH + unit;
                    sv[0] = r; sv[kGruH] = z; sv[2 * kGruH] = n; sv[3 * kGruH] = hn;
                }
            }
            if (SAVE && clip_ok) a.hprev[((size_t)b * a.T_out + t) * (2 * kGruH) + dir * kGruH + unit] = active ? h_own : 0.0f;
            h_own = h_new;
            if (clip_ok) a.out[((size_t)b * a.T_out + t) * (2 * kGruH) + dir * kGruH + unit] = active ? h_new : 0.0f;
            Hn[unit * kGruClips + bl] = h_new;                 // own copy first ...
        }
        __syncthreads();
        // ... then the CTA's contiguous 1 KB slice (32 units x 8 clips) goes to the 7 peers as 128-bit stores
        if (prk != c) {
            const float4* src = reinterpret_cast<const float4*>(Hn + c * kGruUnits * kGruClips);
            float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(Hn, prk) + c * kGruUnits * kGruClips);
            dst[pv] = src[pv];
        }
        cluster.sync();
    }
}

struct GruLayer { float* wih; float* bih; TcWeight tw; float* whh[2]; float* bhh[2]; int din; };

int bigru_recurrence_launch(const GruStepArgs& a, bool save, cudaStream_t st) {
    static cudaError_t attr_rc = cudaFuncSetAttribute(bigru_recurrence_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruSmem);
    static cudaError_t attr_rc2 = cudaFuncSetAttribute(bigru_recurrence_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruSmem);
    AC_CUDA(attr_rc);
    AC_CUDA(attr_rc2);
    AC_REQUIRE(!save || (a.save != nullptr && a.hprev != nullptr), "bigru_recurrence_launch: save buffers missing");
    const int groups = cdiv(a.B, kGruClips);
    AC_TIMED("bigru_recurrence", st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * groups * kGruCluster); cfg.blockDim = dim3(kGruThreads); cfg.dynamicSmemBytes = kGruSmem; cfg.stream = st;
    cudaLaunchAttribute at[1] = {pdl_attr()};
    cfg.attrs = at; cfg.numAttrs = 1;
    if (save) AC_CUDA(cudaLaunchKernelEx(&cfg, bigru_recurrence_kernel<true>, a));
    else AC_CUDA(cudaLaunchKernelEx(&cfg, bigru_recurrence_kernel<false>, a));
    AC_LAUNCHED("bigru_recurrence_kernel");
    return AC_OK;
}

}  // namespace ac

struct ac_bigru {
    float* blob = nullptr;
    int input_dim = 0, layers = 0;
    std::vector<ac::GruLayer> layer;
};

extern "C" {

int ac_bigru_create(const float* const* t, const int64_t* numels, int n_tensors, int input_dim, int hidden, int num_layers,
                    void* stream, ac_bigru_t** out) {
    using namespace ac;
    AC_REQUIRE(t && numels && out, "ac_bigru_create: null argument");
    AC_REQUIRE(hidden == kGruH, "ac_bigru_create: hidden size %d is not supported (the recurrence kernel is built for %d)",
               hidden, kGruH);
    AC_REQUIRE(num_layers >= 1 && input_dim >= 8 && input_dim % 8 == 0, "ac_bigru_create: bad layer count / input size");
    AC_REQUIRE(n_tensors == num_layers * 8, "ac_bigru_create: expected %d tensors (4 per layer and direction), got %d",
               num_layers * 8, n_tensors);
    cudaStream_t st = (cudaStream_t)stream;
    const int H = kGruH;
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += align_up(n, 64); return o; };
    struct Off { size_t wih, bih, pk, whh[2], bhh[2]; };
    std::vector<Off> off(num_layers);
    for (int l = 0; l < num_layers; ++l) {
        const int din = l == 0 ? input_dim : 2 * H;
        for (int d = 0; d < 2; ++d) {
            const int ti = (l * 2 + d) * 4;
            AC_REQUIRE(numels[ti] == (int64_t)3 * H * din && numels[ti + 1] == (int64_t)3 * H * H && numels[ti + 2] == 3 * H &&
                       numels[ti + 3] == 3 * H, "ac_bigru_create: layer %d direction %d has unexpected tensor sizes", l, d);
        }
        off[l].wih = take((size_t)6 * H * din); off[l].bih = take(6 * H); off[l].pk = take(tc_packed_floats(6 * H, din));
        for (int d = 0; d < 2; ++d) { off[l].whh[d] = take((size_t)3 * H * H); off[l].bhh[d] = take(3 * H); }
    }
    ac_bigru_t* net = new ac_bigru_t();
    net->input_dim = input_dim; net->layers = num_layers;
    int rc = check_cuda(cudaMalloc(&net->blob, total * sizeof(float)), "ac_bigru_create: cudaMalloc");
    if (rc != AC_OK) { delete net; return rc; }
    float* B0 = net->blob;
    auto copy = [&](float* dst, const float* src, size_t n) {
        if (rc == AC_OK) rc = check_cuda(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st), "ac_bigru_create copy");
    };
    for (int l = 0; l < num_layers; ++l) {
        const int din = l == 0 ? input_dim : 2 * H;
        GruLayer L{};
        L.din = din; L.wih = B0 + off[l].wih; L.bih = B0 + off[l].bih;
        for (int d = 0; d < 2; ++d) {
            const int ti = (l * 2 + d) * 4;
            copy(L.wih + (size_t)d * 3 * H * din, t[ti], (size_t)3 * H * din);       // rows [d*3H, (d+1)*3H) of the joint projection
            copy(L.bih + d * 3 * H, t[ti + 2], 3 * H);
            L.whh[d] = B0 + off[l].whh[d]; L.bhh[d] = B0 + off[l].bhh[d];
            copy(L.whh[d], t[ti + 1], (size_t)3 * H * H);
            copy(L.bhh[d], t[ti + 3], 3 * H);
        }
        if (rc == AC_OK) rc = tc_pack_weight(L.wih, nullptr, 6 * H, din, B0 + off[l].pk, st, &L.tw);
        net->layer.push_back(L);
    }
    if (rc == AC_OK) rc = check_cuda(cudaStreamSynchronize(st), "ac_bigru_create sync");
    if (rc != AC_OK) { cudaFree(net->blob); delete net; return rc; }
    *out = net;
    return AC_OK;
}

void ac_bigru_destroy(ac_bigru_t* net) {
    if (!net) return;
    cudaFree(net->blob);
    delete net;
}

int ac_bigru_out_dim(const ac_bigru_t* net) { return net ? 2 * ac::kGruH : 0; }

size_t ac_bigru_workspace_bytes(const ac_bigru_t* net, int batch, int T) {
    if (!net) return 0;
    // G [B*T, 6H] + two layer outputs [B, T, 2H]
    return (ac::align_up((size_t)batch * T * 6 * ac::kGruH, 64) + 2 * ac::align_up((size_t)batch * T * 2 * ac::kGruH, 64)) * sizeof(float);
}

int ac_bigru_fwd(const ac_bigru_t* net, const float* x, const int64_t* lens, int B, int T_in, int T_out, float* out,
                 void* workspace, size_t ws_bytes, void* stream) {
    using namespace ac;
    AC_REQUIRE(B >= 0 && T_in >= 1 && T_out >= 0 && T_out <= T_in, "ac_bigru_fwd: bad sizes B=%d T_in=%d T_out=%d", B, T_in, T_out);
    if (B == 0 || T_out == 0) return AC_OK;
    AC_REQUIRE(net && x && lens && out, "ac_bigru_fwd: null argument");
    AC_REQUIRE(workspace && ws_bytes >= ac_bigru_workspace_bytes(net, B, T_in), "ac_bigru_fwd: workspace too small (%zu < %zu)",
               ws_bytes, ac_bigru_workspace_bytes(net, B, T_in));
    cudaStream_t st = (cudaStream_t)stream;
    const int H = kGruH;
    float* G = (float*)workspace;
    float* Y0 = G + align_up((size_t)B * T_in * 6 * H, 64);
    float* Y1 = Y0 + align_up((size_t)B * T_in * 2 * H, 64);
    const float* in = x;
    int t_in = T_in;                       // time stride of the current layer's input
    for (int l = 0; l < net->layers; ++l) {
        const GruLayer& L = net->layer[l];
        GemmArgs g; g.A = in; g.W = L.wih; g.C = G; g.M = B * t_in; g.N = 6 * H; g.K = L.din; g.cbias = L.bih; g.act = ACT_NONE;
        g.tw = &L.tw;
        int rc = gemm_tn(g, st); if (rc) return rc;
        float* y = l + 1 == net->layers ? out : (l & 1 ? Y1 : Y0);
        GruStepArgs a;
        a.G = G; a.lens = lens; a.out = y; a.B = B; a.T_in = t_in; a.T_out = T_out;
        for (int d = 0; d < 2; ++d) { a.whh[d] = L.whh[d]; a.bhh[d] = L.bhh[d]; }
        rc = bigru_recurrence_launch(a, false, st); if (rc) return rc;
        in = y; t_in = T_out;              // the next layer reads [B, T_out, 2H]
    }
    return AC_OK;
}

}  // extern "C"
