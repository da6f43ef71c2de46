// Shared helpers for the audiocaption_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/audiocaption_b200.h"

namespace ac {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return AC_ERR_CUDA;
    }
    return AC_OK;
}

#define AC_CUDA(call)                                               \
    do {                                                            \
        int _rc = ::ac::check_cuda((call), #call);                  \
        if (_rc != AC_OK) return _rc;                               \
    } while (0)

// Optional per-kernel CUDA-event timing (bench.py's roofline leg): when enabled, a pair of events
// brackets every launch on its own stream; ac_timing_report() resolves them after a sync.
int timing_begin(const char* name, cudaStream_t st);
void timing_end(int index, cudaStream_t st);
extern bool g_timing;
struct LaunchTimer {
    cudaStream_t st; bool on; int index = -1;
    LaunchTimer(const char* name, cudaStream_t s) : st(s), on(g_timing) { if (on) index = timing_begin(name, st); }
    ~LaunchTimer() { if (on) timing_end(index, st); }
};
#define AC_TIMED(name, st) ::ac::LaunchTimer _ac_timer_##__LINE__(name, st)

// call after every kernel launch
#define AC_LAUNCHED(name)                                           \
    do {                                                            \
        ::ac::g_launches.fetch_add(1, std::memory_order_relaxed);   \
        int _rc = ::ac::check_cuda(cudaGetLastError(), name);       \
        if (_rc != AC_OK) return _rc;                               \
    } while (0)

#define AC_REQUIRE(cond, ...)                                       \
    do {                                                            \
        if (!(cond)) {                                              \
            ::ac::set_error(__VA_ARGS__);                           \
            return AC_ERR_ARG;                                      \
        }                                                           \
    } while (0)

inline cudaLaunchAttribute pdl_attr() {
    cudaLaunchAttribute a;
    a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a.val.programmaticStreamSerializationAllowed = 1;
    return a;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

#ifdef __CUDACC__
// Programmatic dependent launch: a kernel launched with pdl_attr() may start while its predecessor in the stream
// is still running; everything before pdl_wait() (barrier init, TMEM allocation, weight prefetch, index math)
// overlaps the predecessor's tail, everything after it sees the predecessor's memory.  pdl_trigger() lets the
// NEXT kernel start early in turn.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }
// x * sigmoid(x) on the SFU: ex2.approx (2 ulp) + rcp.approx (1 ulp), ~2^-21 relative error
__device__ __forceinline__ float fast_swish(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// float atomic max valid for any sign (buffer initialised to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
#endif

}  // namespace ac
