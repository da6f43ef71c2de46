// Error handling / bookkeeping of the C ABI.
#include <map>
#include <mutex>

#include "common.cuh"

namespace ac {
static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};
bool g_timing = false;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

struct TimedLaunch { std::string name; cudaEvent_t e0, e1; };
static std::vector<TimedLaunch> g_timed;
static std::mutex g_timed_mu;

int timing_begin(const char* name, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_timed_mu);
    TimedLaunch t; t.name = name;
    cudaEventCreate(&t.e0); cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, st);
    g_timed.push_back(t);
    return (int)g_timed.size() - 1;
}
void timing_end(int index, cudaStream_t st) {     // timers nest (a "span_*" wrapper around per-kernel timers): end by index
    std::lock_guard<std::mutex> lk(g_timed_mu);
    if (index >= 0 && index < (int)g_timed.size()) cudaEventRecord(g_timed[index].e1, st);
}
}  // namespace ac

extern "C" {
int ac_version(void) { return 100; }
const char* ac_last_error(void) { return ac::t_error.c_str(); }
int64_t ac_launch_count(void) { return ac::g_launches.load(); }

void ac_timing_enable(int on) { ac::g_timing = on != 0; }

// Synchronises the device, then writes "name count total_ms\n" lines (one per kernel name) into buf.
int ac_timing_report(char* buf, int buf_len) {
    using namespace ac;
    AC_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_timed_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& t : g_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            auto& a = agg[t.name]; a.first++; a.second += ms;
        } else {
            cudaGetLastError();      // an event that was never recorded: do not leave the error for the next launch check
        }
        cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
    }
    g_timed.clear();
    std::string out;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((int)out.size() + 1 > buf_len) { set_error("ac_timing_report: buffer too small"); return AC_ERR_ARG; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return AC_OK;
}
}
