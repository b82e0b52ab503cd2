// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the sm_100a kernels (gemm_tcgen05.cu, decode_fused.cu).
#pragma once

#include "cap_common.cuh"

namespace cap_ptx {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// ---- fault records -------------------------------------------------------------------------------------------
// A bounded wait that times out traps (reported by CUDA as "unspecified launch failure", after which device memory is
// unreadable).  Before trapping it leaves a record in a pinned, device-mapped HOST buffer (cap_fault_buffer_device in
// cap_core.cu): WHICH wait it was (the source line of the mbar_wait).  The host can read that buffer after the context
// has died (cap_fault_records).  One pointer per translation unit, installed per device.
static __device__ unsigned long long* g_fault_buf = nullptr;

// Deliberately tiny -- one 32-bit store of an immediate (the source line) into the slot line & 63: a record that also
// carried blockIdx / threadIdx cost the chain kernels 36 bytes of register spills at their 128-register cap.
static __device__ __forceinline__ void fault_record_and_trap(int line, uint32_t parity) {
    unsigned long long* buf = g_fault_buf;
    if (buf != nullptr) {
        reinterpret_cast<volatile uint32_t*>(buf)[16 + (line & 63)] = static_cast<uint32_t>(line);   // immediate operands only
        __threadfence_system();
    }
    __trap();
}

// host side: point this translation unit's g_fault_buf at the process-wide buffer, once per device
}  // namespace cap_ptx
extern "C" unsigned long long* cap_fault_buffer_device();
namespace cap_ptx {
static inline int install_fault_buffer() {
    static cap_device_once once;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&once.done, __ATOMIC_ACQUIRE) & bit) return CAP_OK;
    unsigned long long* p = cap_fault_buffer_device();
    if (p != nullptr) CAP_CHECK_CUDA(cudaMemcpyToSymbol(g_fault_buf, &p, sizeof(p)));
    __atomic_fetch_or(&once.done, bit, __ATOMIC_RELEASE);
    return CAP_OK;
}

// Bounded spin: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  The bound is wall time
// (%globaltimer, 10 s), not SM cycles: a CTA can legitimately wait long when the GPU is shared or throttled.
constexpr uint32_t WAIT_LIMIT_TICKS = 10000;   // ticks of 2^20 ns (~1.05 ms): ~10.5 s
__device__ __forceinline__ uint32_t global_ticks() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return static_cast<uint32_t>(t >> 20);
}

__device__ __forceinline__ void mbar_wait_impl(uint64_t* bar, uint32_t parity, int line) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint32_t start = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (spin == 64) start = global_ticks();
        if (spin > 64 && (spin & 1023) == 0 && global_ticks() - start > WAIT_LIMIT_TICKS) fault_record_and_trap(line, parity);
    }
}

// same, acquiring at cluster scope: the arrivals come from the peer CTA of a pair (remote mbarrier.arrive)
__device__ __forceinline__ void mbar_wait_cluster_impl(uint64_t* bar, uint32_t parity, int line) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint32_t start = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (spin == 64) start = global_ticks();
        if (spin > 64 && (spin & 1023) == 0 && global_ticks() - start > WAIT_LIMIT_TICKS) fault_record_and_trap(line, parity);
    }
}
#define mbar_wait(bar, parity) mbar_wait_impl(bar, parity, __LINE__)
#define mbar_wait_cluster(bar, parity) mbar_wait_cluster_impl(bar, parity, __LINE__)

// arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                            int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner),
        "r"(c_outer)
        : "memory");
}

// L2 eviction policies: weights are re-read by every row tile of every batch in flight (keep: evict_last);
// K|V rows, caches and logits are streamed once per step (evict_first) and must not push the weights out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                                 int c_outer, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row groups are 1024 bytes apart (SBO),
// LBO unused for swizzled K-major, descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(const void* tile) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_u32(tile) & 0x3FFFF) >> 4);  // [0,14)  start address
    desc |= static_cast<uint64_t>(1) << 16;                          // [16,30) leading byte offset (ignored)
    desc |= static_cast<uint64_t>(1024 >> 4) << 32;                  // [32,46) stride byte offset
    desc |= static_cast<uint64_t>(1) << 46;                          // [46,48) version
    desc |= static_cast<uint64_t>(2) << 61;                          // [61,64) SWIZZLE_128B
    return desc;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, shape M x N.
__host__ __device__ constexpr uint32_t make_instr_desc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- cta_group::2 forms: a pair of CTAs (one cluster, adjacent SMs) works on one 256-row tile ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Both CTAs load into their own smem but signal the LEADER's mbarrier (peer bit of the shared-window address
// cleared), so one barrier phase covers the four tiles of a stage.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                                int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                                     int c_outer, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer), "l"(policy)
        : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// read a float at the same shared-memory offset in CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// registers -> TMEM, same 32 lanes x 32 columns shape (used to park residual + projection between the two
// LayerNorm passes); tmem_st_wait() before the values are read back
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, naming the registers of an earlier tcgen05.ld as in/out operands: whatever consumes them is then
// data-dependent on the wait (needed when another load is issued between a load and the use of its values).
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}


// Shared-memory matrix descriptor for a K-major operand stored WITHOUT swizzle as 8-row x 16-byte core
// matrices: `lbo` = byte distance between core matrices adjacent along K, `sbo` = byte distance between
// 8-row groups along M/N (descriptor version 1, layout type 0).
__device__ __forceinline__ uint64_t make_smem_desc_noswizzle(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    desc |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    desc |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    return desc;
}

// One lane of the (converged) warp is elected.  tcgen05.mma / tcgen05.commit / TMA are uniform-datapath
// instructions: guarded by elect.sync, ptxas emits them once with uniform registers.  Guarded by
// `if (lane == 0)` instead, it wraps every one of them in an ELECT / R2UR / BRA.U.ANY loop over the active
// lanes, and a single thread then needs ~200 cycles per MMA issue (measured, profiles/README.md).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "elect.sync %%rx|%%px, %1;\n"
        "@%%px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// generic-proxy writes (st.shared / st.global) made visible to the async proxy (tcgen05.mma operand reads, TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// same for shared memory only (the unqualified form costs a MEMBAR.ALL.GPU)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 1-D bulk copy global -> shared (bytes multiple of 16, both 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace cap_ptx
