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

// ---- fault records (see tcgen05_ptx.cuh) -----------------------------------------------------------------------
// 64 bytes of header + 64 32-bit records (slot = source line & 63, value = source line), in pinned host memory that
// every device of the process can write (portable + mapped); readable after a sticky CUDA error.
namespace {
unsigned long long* g_fault_host = nullptr;
}

extern "C" unsigned long long* cap_fault_buffer_device() {
    static const bool ok = [] {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 4096, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        memset(p, 0, 4096);
        g_fault_host = static_cast<unsigned long long*>(p);
        return true;
    }();
    if (!ok) return nullptr;
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, g_fault_host, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<unsigned long long*>(d);
}

// Copies the non-empty fault records (source lines of the waits that timed out) into out, at most max_records, and
// returns their number (0 = no wait ever timed out).  Works after the CUDA context has faulted.
extern "C" int cap_fault_records(unsigned int* out, int max_records) {
    if (g_fault_host == nullptr) return 0;
    const unsigned int* slots = reinterpret_cast<const unsigned int*>(g_fault_host) + 16;
    int n = 0;
    for (int i = 0; i < 64; ++i) {
        const unsigned int r = __atomic_load_n(slots + i, __ATOMIC_ACQUIRE);
        if (r == 0) continue;
        if (n < max_records) out[n] = r;
        ++n;
    }
    return n;
}

// ---- flight recorder (cap_common.cuh) -----------------------------------------------------------------------------
// u32 counters [kind][entered, left] behind the fault records in the same pinned page; off unless OPENVIIC_FLIGHT=1.
extern "C" unsigned int* cap_flight_buffer_device() {
    static const bool on = getenv("OPENVIIC_FLIGHT") && atoi(getenv("OPENVIIC_FLIGHT")) != 0;
    if (!on) return nullptr;
    unsigned long long* base = cap_fault_buffer_device();
    return base ? reinterpret_cast<unsigned int*>(base) + 256 : nullptr;
}

// Copies the 2 * kinds counters ([kind][entered, left]) into out and returns the number of kinds (0: recorder off).
extern "C" int cap_flight_records(unsigned int* out, int max_kinds) {
    if (g_fault_host == nullptr || cap_flight_buffer_device() == nullptr) return 0;
    const unsigned int* f = reinterpret_cast<const unsigned int*>(g_fault_host) + 256;
    const int kinds = max_kinds < 16 ? max_kinds : 16;
    for (int i = 0; i < 2 * kinds; ++i) out[i] = __atomic_load_n(f + i, __ATOMIC_ACQUIRE);
    return kinds;
}
