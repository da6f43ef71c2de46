// Error handling / bookkeeping of the C ABI.
#include "common.cuh"

namespace ac {
static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}
}  // namespace ac

extern "C" {
int ac_version(void) { return 100; }
const char* ac_last_error(void) { return ac::t_error.c_str(); }
int64_t ac_launch_count(void) { return ac::g_launches.load(); }
}
