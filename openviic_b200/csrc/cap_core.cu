// Error reporting, ABI version and launch accounting shared by every entry point.
#include "cap_common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

std::atomic<long long> g_cap_launches{0};

namespace {
thread_local char g_last_error[1024] = "";
}

int cap_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

bool cap_pdl_enabled() {
    static const bool on = [] {
        const char* v = getenv("OPENVIIC_PDL");
        return !(v && v[0] == '0');
    }();
    return on;
}

extern "C" int cap_abi_version(void) { return CAP_ABI_VERSION; }

extern "C" const char* cap_last_error(void) { return g_last_error; }

extern "C" int64_t cap_launch_count(void) { return static_cast<int64_t>(g_cap_launches.load()); }
