// Spatial partition of the GPU's SMs into two sets with one stream each (CUDA green contexts, driver API; resolved through
// the runtime's driver entry points, so there is no link-time libcuda dependency).
//
// Why: one optimizer step of the captioner is a tensor-core-bound frozen encoder (Cnn14: persistent 148-CTA convolution
// kernels) followed by ~250 latency-bound launches of a few CTAs each (bi-GRU, decoder, backward passes).  The encoder of
// batch i+1 does not depend on step i, so the two can run side by side -- but on ordinary streams the block scheduler hands
// every SM a convolution CTA frees straight to the next (already pending) convolution CTA, and the small kernels (several
// need 64 SMs at once as 8-CTA clusters) starve: measured, both halves stretch to the sum of their durations.  With disjoint
// SM sets neither can take the other's SMs.  audiocaption_b200/train_step.py `TrainStep.prefetch` is the user.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {
template <typename F>
F driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(p);
}
using FnGetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
using FnSplit = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
using FnDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
using FnCtxCreate = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
using FnCtxDestroy = CUresult (*)(CUgreenCtx);
using FnStreamCreate = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
using FnDeviceGet = CUresult (*)(CUdevice*, int);
}  // namespace

struct ac_sm_partition {   // == ac_sm_partition_t of the header
    CUgreenCtx ctx[2] = {nullptr, nullptr};
    CUstream stream[2] = {nullptr, nullptr};
    int sms[2] = {0, 0};
};

extern "C" {

// Splits the current device's SMs into set 0 (about `sms_first` SMs, rounded to the hardware's 8-SM granularity) and set 1
// (the other whole groups); one non-blocking stream per set (set 1's at high priority).  Fails loudly (AC_ERR_CUDA + message)
// when the driver has no green contexts; the caller then keeps ordinary streams.
int ac_sm_partition_create(int sms_first, ac_sm_partition** out) {
    using namespace ac;
    AC_REQUIRE(out != nullptr && sms_first >= 8, "ac_sm_partition_create: bad argument");
    auto get_res = driver_fn<FnGetRes>("cuDeviceGetDevResource");
    auto split = driver_fn<FnSplit>("cuDevSmResourceSplitByCount");
    auto gen_desc = driver_fn<FnDesc>("cuDevResourceGenerateDesc");
    auto ctx_create = driver_fn<FnCtxCreate>("cuGreenCtxCreate");
    auto ctx_destroy = driver_fn<FnCtxDestroy>("cuGreenCtxDestroy");
    auto stream_create = driver_fn<FnStreamCreate>("cuGreenCtxStreamCreate");
    auto device_get = driver_fn<FnDeviceGet>("cuDeviceGet");
    AC_REQUIRE(get_res && split && gen_desc && ctx_create && ctx_destroy && stream_create && device_get,
               "ac_sm_partition_create: this driver has no green-context API");
    int ordinal = 0;
    AC_CUDA(cudaGetDevice(&ordinal));
    AC_CUDA(cudaFree(nullptr));                                   // the primary context exists
    CUdevice dev;
    CUresult cr = device_get(&dev, ordinal);
    AC_REQUIRE(cr == CUDA_SUCCESS, "ac_sm_partition_create: cuDeviceGet failed (%d)", (int)cr);
    CUdevResource all;
    cr = get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
    AC_REQUIRE(cr == CUDA_SUCCESS, "ac_sm_partition_create: cuDeviceGetDevResource failed (%d)", (int)cr);
    // whole 8-SM groups (the granularity that keeps 8-CTA clusters schedulable inside a set), then two descriptors over them
    unsigned int n_groups = 0;
    cr = split(nullptr, &n_groups, &all, nullptr, 0, 8);
    AC_REQUIRE(cr == CUDA_SUCCESS && n_groups >= 2, "ac_sm_partition_create: cannot split %u SMs into 8-SM groups (%d, %u groups)",
               all.sm.smCount, (int)cr, n_groups);
    std::vector<CUdevResource> groups(n_groups);
    CUdevResource rest;
    cr = split(groups.data(), &n_groups, &all, &rest, 0, 8);
    AC_REQUIRE(cr == CUDA_SUCCESS && n_groups >= 2, "ac_sm_partition_create: cuDevSmResourceSplitByCount failed (%d)", (int)cr);
    // The 8-SM groups do not cover the chip (B200: 15 groups = 120 of 148 SMs; GPCs hold 16-20 SMs): the remainder goes to
    // set 0, whose user (the convolutions) needs no clusters.  Set 0 = remainder + as many groups as `sms_first` asks for.
    const int rest_sms = (int)rest.sm.smCount;
    unsigned int first = std::min<unsigned int>(n_groups - 1, (unsigned int)std::max(1, (sms_first - rest_sms + 4) / 8));
    ac_sm_partition* p = new ac_sm_partition();
    for (int s = 0; s < 2 && cr == CUDA_SUCCESS; ++s) {
        std::vector<CUdevResource> set(groups.begin() + (s == 0 ? 0 : first), groups.begin() + (s == 0 ? first : n_groups));
        if (s == 0 && rest_sms > 0) set.push_back(rest);
        CUdevResource* g0 = set.data();
        const unsigned int n = (unsigned int)set.size();
        for (unsigned int i = 0; i < n; ++i) p->sms[s] += (int)g0[i].sm.smCount;
        CUdevResourceDesc desc;
        cr = gen_desc(&desc, g0, n);
        if (cr == CUDA_SUCCESS) cr = ctx_create(&p->ctx[s], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
        if (cr == CUDA_SUCCESS) cr = stream_create(&p->stream[s], p->ctx[s], CU_STREAM_NON_BLOCKING, s == 1 ? -1 : 0);
    }
    if (cr != CUDA_SUCCESS) {
        set_error("ac_sm_partition_create: green context setup failed (%d)", (int)cr);
        for (int s = 0; s < 2; ++s) {
            if (p->stream[s]) cudaStreamDestroy((cudaStream_t)p->stream[s]);
            if (p->ctx[s]) ctx_destroy(p->ctx[s]);
        }
        delete p;
        return AC_ERR_CUDA;
    }
    *out = p;
    return AC_OK;
}

void* ac_sm_partition_stream(const ac_sm_partition* p, int which) { return p && (which == 0 || which == 1) ? (void*)p->stream[which] : nullptr; }
int ac_sm_partition_sms(const ac_sm_partition* p, int which) { return p && (which == 0 || which == 1) ? p->sms[which] : 0; }

void ac_sm_partition_destroy(ac_sm_partition* p) {
    if (!p) return;
    auto ctx_destroy = driver_fn<FnCtxDestroy>("cuGreenCtxDestroy");
    for (int s = 0; s < 2; ++s) {
        if (p->stream[s]) { cudaStreamSynchronize((cudaStream_t)p->stream[s]); cudaStreamDestroy((cudaStream_t)p->stream[s]); }
        if (p->ctx[s] && ctx_destroy) ctx_destroy(p->ctx[s]);
    }
    delete p;
}

}  // extern "C"
