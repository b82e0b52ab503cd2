// Shared device/host helpers for the caption hot-path kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <utility>

#include "../../include/openviic_cap.h"

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---------------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry point returns an int and records a message (no exceptions,
// no exit() across the ABI -- SURVEY.md section 8b "Error convention").
// ---------------------------------------------------------------------------------------------
int cap_set_error(int code, const char* fmt, ...);

#define CAP_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return cap_set_error(CAP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define CAP_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) return cap_set_error(CAP_ERR_INVALID, __VA_ARGS__);                       \
    } while (0)

#define CAP_PROPAGATE(expr)                                                                    \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc != CAP_OK) return _rc;                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic Dependent Launch: every kernel is launched with the programmatic-stream-serialization
// attribute, signals `launch_dependents` on entry and blocks in `griddepcontrol.wait` before it touches
// memory a predecessor may still be writing.  The next kernel's launch latency and prologue (barrier
// init, TMEM allocation, descriptor prefetch, bias staging) thereby overlap the current kernel's tail
// -- in eager streams and inside captured CUDA graphs alike.  OPENVIIC_PDL=0 turns the attribute off
// (the device instructions are then no-ops).
// ---------------------------------------------------------------------------------------------
bool cap_pdl_enabled();
// pinned, device-mapped host words for the flight recorder (cap_core.cu); nullptr unless OPENVIIC_FLIGHT=1
extern "C" unsigned int* cap_flight_buffer_device();

// "Already done on this device" flag: one atomic bit per device ordinal (used for per-device kernel attributes and
// per-translation-unit device symbols; setting either twice is harmless).
struct cap_device_once {
    unsigned long long done = 0;
};

#ifdef __CUDACC__
// ---- flight recorder (debug, OPENVIIC_FLIGHT=1) -----------------------------------------------------------------
// One thread per CTA counts "entered" / "left" per kernel kind in pinned, device-mapped HOST memory (posted
// reductions, nobody waits for them).  When the device hangs the host reads which kinds have CTAs that entered and
// never left (cap_flight_records) -- without the recorder a hang in a wait that is not bounded (griddepcontrol.wait,
// a cluster barrier, tcgen05.alloc) says nothing about where it is.  One pointer per translation unit; nullptr
// (the default) costs one uniform branch per CTA.
static __device__ unsigned int* g_flight = nullptr;
// point this translation unit's g_flight at the process-wide buffer, once per device (no-op when the recorder is off)
static inline void cap_install_flight() {
    static cap_device_once once;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&once.done, __ATOMIC_ACQUIRE) & bit) return;
    unsigned int* p = cap_flight_buffer_device();
    if (p != nullptr) cudaMemcpyToSymbol(g_flight, &p, sizeof(p));
    __atomic_fetch_or(&once.done, bit, __ATOMIC_RELEASE);
}
#endif

template <bool PDL = true, typename Kernel, typename... Args>
inline void cap_launch_kernel(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
#ifdef __CUDACC__
    cap_install_flight();
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (PDL && cap_pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {  // thread-block cluster along x (CTA pairs of the 2-CTA GEMM)
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define CAP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cap_launch_kernel(kernel, dim3(grid), dim3(block), smem, stream, 1, __VA_ARGS__)
// Full stream serialisation (no programmatic dependent launch): the kernel starts only after everything before it on
// the stream has COMPLETED.  Used as the fence between phases whose later kernels prefetch data before their own
// griddepcontrol.wait (the decode cross-attention fetches the encode-time K|V early).
#define CAP_LAUNCH_SERIAL(kernel, grid, block, smem, stream, ...) \
    cap_launch_kernel<false>(kernel, dim3(grid), dim3(block), smem, stream, 1, __VA_ARGS__)

// Opt-in to > 48 KB of dynamic shared memory: the attribute is PER DEVICE, and entry points may be called from several
// host threads.
template <typename Kernel>
inline int cap_opt_in_smem(cap_device_once& once, Kernel kernel, int bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&once.done, __ATOMIC_ACQUIRE) & bit) return CAP_OK;
    CAP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    __atomic_fetch_or(&once.done, bit, __ATOMIC_RELEASE);
    return CAP_OK;
}

static inline int cap_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cap_set_error(CAP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return CAP_OK;
}

// ---------------------------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

enum FlightKind {
    FK_CHAIN = 0, FK_CHAIN_READY = 1 /* past TMEM allocation and the first cluster barrier */, FK_CROSS_PRODUCER = 2,
    FK_CROSS_CONSUMER = 3, FK_CROSS_CONSUMER_READY = 4 /* past griddepcontrol.wait */, FK_SELF_ATTENTION = 5, FK_GEMM = 6,
    FK_GEMM_READY = 7, FK_ATTENTION = 8, FK_BEAM_MERGE = 9, FK_BEAM_SELECT = 10, FK_BEAM_OTHER = 11, FK_ROW_KERNELS = 12,
    FK_KINDS = 16
};
__device__ __forceinline__ void flight_mark(int kind, int left) {
    unsigned int* f = g_flight;
    if (f != nullptr) asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(f + 2 * kind + left) : "memory");
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// simple kernels: let the successor start launching, then wait for the predecessor's results
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}

// 2^x on the SFU, one instruction (results below 2^-126 flush to zero).  exp2f() wraps the same MUFU.EX2 in a range
// test and two multiplies for denormal results: four instructions per element in the softmax / log-sum-exp loops,
// which are issue-bound.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

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

// 8 bf16 <-> 8 floats through one 16-byte vector
struct __align__(16) bf16x8 {
    bf162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(p.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ bf16x8 pack8(const float* f) {
    bf16x8 p;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return p;
}

// Ordering used everywhere a "stable descending sort" is restated: larger value first, and on
// equal values the smaller flat index first (the reference's torch.sort tie order on its CPU
// path, models/modules/beam_search.py:37).
__device__ __forceinline__ bool cand_before(float va, int ia, float vb, int ib) {
    return (va > vb) || (va == vb && ia < ib);
}

#endif  // __CUDACC__
