// y[M,N] = act(x[M,K] . w[N,K]^T + bias) -- bf16 operands, fp32 accumulation in TMEM.
//
// Blackwell-native structure (one output tile per CTA, 128 x BLOCK_N):
//   warp 0  : TMA producer   (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 bf16 along K)
//   warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16)
//   warps 2-5: epilogue       (tcgen05.ld 32x32b -> bias/activation -> vectorised global store)
// smem ring of `num_stages` {A 128x64, B BLOCK_Nx64} tiles guarded by full/empty mbarriers;
// tcgen05.commit releases a stage back to the producer and finally signals the epilogue.
//
// Replaces every nn.Linear on the caption path (see include/openviic_cap.h: cap_linear).
#include "cap_common.cuh"
#include "tcgen05_ptx.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

extern std::atomic<long long> g_cap_launches;

namespace {

using namespace cap_ptx;
constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quadrant, each drains half the columns
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int MAX_STAGES = 8;
constexpr uint32_t A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;

struct GemmParams {
    void* out;
    const float* bias;
    int M, N, K, ldo;
    int out_f32, act, num_stages, vec_ok;
    uint32_t pipe_bytes;  // bytes reserved for the stage ring (>= the epilogue's staging tile), multiple of 1024
    float* part_ms;       // STATS epilogue: [M][chunks][2] (max, sum exp(x - max)) per 32-column chunk
    // LN epilogue (cluster of N/128 CTAs along N): out = LayerNorm(residual + x.w^T + bias) * gamma + beta (+pos)
    const float* ln_gamma;
    const float* ln_beta;
    const float* residual;   // fp32 [M][ldr] or nullptr
    const float* pos;        // fp32 [pos_rows][N] or nullptr
    const uint8_t* zero_rows;  // uint8 [M] or nullptr: rows written as zeros
    float* out32;            // fp32 copy [M][ldo32] or nullptr (`out` is the bf16 copy or nullptr)
    int ldr, ldo32, pos_rows;
    float eps;
    unsigned long long* trace;  // debug: 8 %globaltimer stamps per CTA (cap_debug_gemm_trace), else nullptr
    // split-K (cap_linear_splitk): blockIdx.z takes k-blocks [z * kb_per_split, ...) and writes its partial product to
    // out + z * split_stride_bytes; 0 = the whole K in one CTA
    int kb_per_split;
    size_t split_stride_bytes;
};

__device__ __forceinline__ void stamp(const GemmParams& p, int slot) {
    if (p.trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + slot] = t;
    }
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == CAP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == CAP_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == CAP_ACT_LEAKY_RELU) return v > 0.f ? v : 0.01f * v;   // F.leaky_relu default slope (encoders.py:243-245)
    return v;
}

// ------------------------------------------------------------------------------------------ kernel
// STATS (vocabulary projection, fp32 logits out): besides the store, every row emits (max, sum exp(x - max))
// for each 32-column chunk.  beam_rowmerge_kernel (beam.cu) turns those into the row's log-sum-exp and
// reads only the `beam` chunks with the largest maxima -- the row's top-`beam` logits provably lie there --
// so the 52 MB logits buffer is written once and almost never read back.
//
// LN (BLOCK_N = 128, 1-CTA MMA): the N/128 CTAs of one row tile form a cluster along N; the epilogue adds
// bias and the fp32 residual, keeps the tile in shared memory, exchanges per-row sums and centred sums of
// squares with its cluster peers through distributed shared memory (two cluster barriers), normalises and
// writes a bf16 copy (next GEMM operand) and an fp32 copy (next residual).  One kernel instead of
// GEMM -> fp32 round trip -> LayerNorm kernel.
template <int BLOCK_N, bool STATS, bool LN = false>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_tn_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const GemmParams p) {
    if (threadIdx.x == 0) flight_mark(FK_GEMM, 0);
    constexpr uint32_t B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

    // alignment from the declaration, not from integer arithmetic (which would turn every shared-memory access
    // of the epilogue into a generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const int stages = p.num_stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.pipe_bytes);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y;   // grid (n tiles, m tiles)
    const int n_tile = blockIdx.x;
    const int n_tiles = gridDim.x;
    const int m0 = m_tile * BLOCK_M;
    const int n0 = n_tile * BLOCK_N;
    const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int kb_begin = p.kb_per_split > 0 ? static_cast<int>(blockIdx.z) * p.kb_per_split : 0;
    const int num_kb = p.kb_per_split > 0 ? min(p.kb_per_split, num_kb_total - kb_begin) : num_kb_total;

    if (threadIdx.x == 0) stamp(p, 0);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(tmem_full_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(p, 1);
    if (threadIdx.x == 0) flight_mark(FK_GEMM_READY, 0);
    // The next kernel may start its prologue now -- only NOW, with this CTA's Tensor Memory in hand: a dependent that
    // starts earlier can allocate Tensor Memory on this SM and then park in griddepcontrol.wait for this grid, while
    // this CTA blocks in tcgen05.alloc behind it (no cycle is possible once every CTA of the primary holds what it needs).
    pdl_launch_dependents();

    // Producer and MMA warps run their loops with the WHOLE warp (uniform control flow) and elect one lane for
    // the TMA / tcgen05 instructions: issued from an `if (lane == 0)` region every tcgen05.mma was wrapped in an
    // ELECT / R2UR / BRA.U.ANY loop and cost the single thread ~200 cycles (3x the MMA's own 64-cycle floor).
    if (warp == 0) {
        const uint64_t keep_policy = l2_policy_evict_last();  // weights stay in L2 across the batches in flight
        pdl_wait();  // A (and only A) may still be in flight from the previous kernel
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % stages;
            const uint32_t phase = (kb / stages) & 1;
            mbar_wait(&empty_bar[s], phase ^ 1);
            uint8_t* a_tile = smem + s * STAGE_BYTES;
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(a_tile, &tmap_a, &full_bar[s], (kb_begin + kb) * BLOCK_K, m0);
                tma_load_2d_hint(a_tile + A_TILE_BYTES, &tmap_b, &full_bar[s], (kb_begin + kb) * BLOCK_K, n0, keep_policy);  // weights
                if (kb == 0) stamp(p, 2);
            }
            __syncwarp();
        }
        if constexpr (LN) { __syncwarp(); cluster_sync(); cluster_sync(); }  // the epilogue's two exchanges
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = make_instr_desc(BLOCK_M, BLOCK_N);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t phase = (kb / stages) & 1;
                mbar_wait(&full_bar[s], phase);
                tcgen05_fence_after();
                const uint8_t* a_tile = smem + s * STAGE_BYTES;
                const uint64_t a_desc = make_smem_desc(a_tile);
                const uint64_t b_desc = make_smem_desc(a_tile + A_TILE_BYTES);
                if (elect_one_sync()) {
                    if (kb == 0) stamp(p, 3);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                        umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);   // stage reusable once these MMAs have read it
                    if (kb == num_kb - 1) {       // accumulator complete
                        umma_commit(tmem_full_bar);
                        stamp(p, 4);
                    }
                }
                __syncwarp();
            }
        }
        if constexpr (LN) { __syncwarp(); cluster_sync(); cluster_sync(); }
    } else {
        // ---- epilogue: TMEM -> registers -> (+bias, activation) -> smem transpose -> coalesced rows ----
        const int quad = warp & 3;            // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;     // which half of each column pass this warp drains (0 or 1)
        const int etid = threadIdx.x - 64;    // 0..255 among the epilogue threads
        for (int i = etid; i < BLOCK_N; i += EPI_WARPS * 32)
            s_bias[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
        float* s_gamma = s_bias + BLOCK_N;          // LN only: gamma | beta | row sums [2][128] | row sq [2][128]
        float* s_beta = s_gamma + BLOCK_N;
        float* s_sum = s_beta + BLOCK_N;
        float* s_sq = s_sum + 2 * BLOCK_M;
        if constexpr (LN) {
            for (int i = etid; i < BLOCK_N; i += EPI_WARPS * 32) {
                s_gamma[i] = __ldg(p.ln_gamma + n0 + i);
                s_beta[i] = __ldg(p.ln_beta + n0 + i);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");  // bias tile visible to the epilogue warps
        mbar_wait(tmem_full_bar, 0);
        if (warp == 2 && lane == 0) stamp(p, 5);
        tcgen05_fence_after();
        if constexpr (LN) {
            pdl_wait();  // residual / zero_rows were written by earlier kernels
            constexpr int PITCH = BLOCK_N * 4 + 16;
            const int trow = quad * 32 + lane;
            const int grow = m0 + trow;
            const bool row_ok = grow < p.M;
            const uint32_t cluster_n = static_cast<uint32_t>(n_tiles);
            float* my_row = reinterpret_cast<float*>(smem + static_cast<size_t>(trow) * PITCH);
            const float* rrow = (p.residual != nullptr && row_ok) ? p.residual + static_cast<size_t>(grow) * p.ldr + n0 : nullptr;
            float sum = 0.f;
#pragma unroll 1
            for (int c0 = 32 * half; c0 < BLOCK_N; c0 += 64) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rrow) r4 = *reinterpret_cast<const float4*>(rrow + c0 + j);
                    float4 y;
                    y.x = __uint_as_float(v[j]) + s_bias[c0 + j] + r4.x;
                    y.y = __uint_as_float(v[j + 1]) + s_bias[c0 + j + 1] + r4.y;
                    y.z = __uint_as_float(v[j + 2]) + s_bias[c0 + j + 2] + r4.z;
                    y.w = __uint_as_float(v[j + 3]) + s_bias[c0 + j + 3] + r4.w;
                    sum += (y.x + y.y) + (y.z + y.w);
                    *reinterpret_cast<float4*>(my_row + c0 + j) = y;
                }
            }
            s_sum[half * BLOCK_M + trow] = sum;
            cluster_sync();  // #1: every CTA of the row tile has published its partial row sums
            float tot = 0.f;
            for (uint32_t rk = 0; rk < cluster_n; ++rk)
                tot += ld_dsmem_f32(s_sum + trow, rk) + ld_dsmem_f32(s_sum + BLOCK_M + trow, rk);
            const float mean = tot / static_cast<float>(p.N);
            float sq = 0.f;
#pragma unroll 1
            for (int c0 = 32 * half; c0 < BLOCK_N; c0 += 64) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 y = *reinterpret_cast<const float4*>(my_row + c0 + j);
                    const float a0 = y.x - mean, a1 = y.y - mean, a2 = y.z - mean, a3 = y.w - mean;
                    sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
                }
            }
            s_sq[half * BLOCK_M + trow] = sq;
            cluster_sync();  // #2
            float vs = 0.f;
            for (uint32_t rk = 0; rk < cluster_n; ++rk)
                vs += ld_dsmem_f32(s_sq + trow, rk) + ld_dsmem_f32(s_sq + BLOCK_M + trow, rk);
            const float rstd = rsqrtf(vs / static_cast<float>(p.N) + p.eps);
            const bool zero = p.zero_rows != nullptr && row_ok && p.zero_rows[grow] != 0;
            const float* prow = (p.pos != nullptr && row_ok)
                                    ? p.pos + static_cast<size_t>(grow % p.pos_rows) * p.N + n0 : nullptr;
#pragma unroll 1
            for (int c0 = 32 * half; c0 < BLOCK_N; c0 += 64) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 y = *reinterpret_cast<const float4*>(my_row + c0 + j);
                    float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (prow) pp = *reinterpret_cast<const float4*>(prow + c0 + j);
                    y.x = (y.x - mean) * rstd * s_gamma[c0 + j] + s_beta[c0 + j] + pp.x;
                    y.y = (y.y - mean) * rstd * s_gamma[c0 + j + 1] + s_beta[c0 + j + 1] + pp.y;
                    y.z = (y.z - mean) * rstd * s_gamma[c0 + j + 2] + s_beta[c0 + j + 2] + pp.z;
                    y.w = (y.w - mean) * rstd * s_gamma[c0 + j + 3] + s_beta[c0 + j + 3] + pp.w;
                    if (zero) y = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(my_row + c0 + j) = y;
                }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");  // both halves of the quadrant normalised
            // coalesced write-out of the quadrant's 32 rows (16 per warp): 4 columns per lane-vector
            constexpr int VECS = BLOCK_N / 4;
            for (int idx = half * 16 * VECS + lane; idx < (half + 1) * 16 * VECS; idx += 32) {
                const int rr = idx / VECS, vv = idx % VECS;
                const int orow = m0 + quad * 32 + rr;
                if (orow >= p.M) continue;
                const float4 y = *reinterpret_cast<const float4*>(smem + static_cast<size_t>(quad * 32 + rr) * PITCH + vv * 16);
                const size_t col = static_cast<size_t>(n0) + vv * 4;
                if (p.out32 != nullptr) *reinterpret_cast<float4*>(p.out32 + static_cast<size_t>(orow) * p.ldo32 + col) = y;
                if (p.out != nullptr) {
                    uint2 packed;
                    packed.x = *reinterpret_cast<const uint32_t*>(&static_cast<const bf162&>(__floats2bfloat162_rn(y.x, y.y)));
                    packed.y = *reinterpret_cast<const uint32_t*>(&static_cast<const bf162&>(__floats2bfloat162_rn(y.z, y.w)));
                    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(orow) * p.ldo + col) = packed;
                }
            }
        } else {
        // All MMAs have retired, so the pipeline stages are free: reuse them as the output staging tile.
        // Tiles wider than 128 columns go through the staging buffer in 128-column passes.
        constexpr int EPI_N = BLOCK_N < 128 ? BLOCK_N : 128;
        const int esz = p.out_f32 ? 4 : 2;
        const int row_bytes = EPI_N * esz;
        const int pitch = row_bytes + 16;     // +16 B: consecutive rows start one bank group apart
        uint8_t* stage_out = smem + static_cast<size_t>(quad * 32) * pitch;
        uint8_t* my_row = stage_out + static_cast<size_t>(lane) * pitch;
#pragma unroll 1
        for (int h0 = 0; h0 < BLOCK_N; h0 += EPI_N) {
        // the two warps of a quadrant split the pass: 32-column chunks alternate between them
#pragma unroll 1
        for (int c0 = h0 + 32 * half; c0 < h0 + EPI_N; c0 += 64) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(__uint_as_float(v[j]) + s_bias[c0 + j], p.act);
            if constexpr (STATS) {
                const int grow = m0 + quad * 32 + lane;
                const int valid = p.N - (n0 + c0);  // columns of this chunk inside the vocabulary (warp-uniform)
                float cm = -INFINITY, cs = 0.f;
                if (valid >= 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, f[j]);
                    const float cm2 = cm * 1.4426950408889634f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) cs += fast_exp2(fmaf(f[j], 1.4426950408889634f, -cm2));
                } else if (valid > 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, j < valid ? f[j] : -INFINITY);
#pragma unroll
                    for (int j = 0; j < 32; ++j) cs += j < valid ? __expf(f[j] - cm) : 0.f;
                }
                if (grow < p.M) {
                    const size_t chunk = static_cast<size_t>(n_tile) * (BLOCK_N / 32) + c0 / 32;
                    const size_t chunks = static_cast<size_t>(n_tiles) * (BLOCK_N / 32);
                    *reinterpret_cast<float2*>(p.part_ms + (static_cast<size_t>(grow) * chunks + chunk) * 2) =
                        make_float2(cm, cs);
                }
            }
            if (p.out_f32) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(my_row + (c0 - h0 + j) * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8) *reinterpret_cast<bf16x8*>(my_row + (c0 - h0 + j) * 2) = pack8(f + j);
            }
        }
        // both warps of the quadrant have staged their chunks (named barrier 2+quad, 64 threads)
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
        // write-out: the quadrant's 32 rows are split between its two warps; lanes tile a row with 16-byte vectors
        const int vecs_per_row = row_bytes / 16;
        const int elems_per_vec = 16 / esz;
        const int n_valid = min(EPI_N, p.N - (n0 + h0));            // valid columns of this pass
        uint8_t* gout = reinterpret_cast<uint8_t*>(p.out) + static_cast<size_t>(blockIdx.z) * p.split_stride_bytes;
        for (int idx = half * 16 * vecs_per_row + lane; idx < (half + 1) * 16 * vecs_per_row; idx += 32) {
            const int rr = idx / vecs_per_row, vv = idx % vecs_per_row;
            const int grow = m0 + quad * 32 + rr;
            const int col = vv * elems_per_vec;
            if (grow >= p.M || col >= n_valid) continue;
            const uint8_t* src = stage_out + static_cast<size_t>(rr) * pitch + vv * 16;
            uint8_t* dst = gout + (static_cast<size_t>(grow) * p.ldo + n0 + h0 + col) * esz;
            if (p.vec_ok && col + elems_per_vec <= n_valid) {
                *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(src);
            } else {  // ragged last tile or unaligned output: element-wise
                const int cnt = min(elems_per_vec, n_valid - col);
                if (p.out_f32) {
                    for (int e = 0; e < cnt; ++e) reinterpret_cast<float*>(dst)[e] = reinterpret_cast<const float*>(src)[e];
                } else {
                    for (int e = 0; e < cnt; ++e) reinterpret_cast<bf16*>(dst)[e] = reinterpret_cast<const bf16*>(src)[e];
                }
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");  // the next pass overwrites the staging rows
        }  // 128-column passes
        }
    }
    if (warp == 2 && lane == 0) stamp(p, 6);
    tcgen05_fence_before();
    // 2-CTA: both epilogues done before TMEM is released; LN: no CTA leaves while peers may read its statistics
    if constexpr (LN) cluster_sync(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
    if (threadIdx.x == 32) stamp(p, 7);
    if (threadIdx.x == 0) {
        flight_mark(FK_GEMM, 1);
        flight_mark(FK_GEMM_READY, 1);
    }
}

// ---------------------------------------------------------------------------------- host launcher
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride `ld` elements; box = box_rows x 64, SWIZZLE_128B.
int make_tmap(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return cap_set_error(CAP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return cap_set_error(CAP_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%d cols=%d ld=%d",
                             static_cast<int>(r), rows, cols, ld);
    return CAP_OK;
}

template <int BLOCK_N, bool STATS = false, bool LN = false>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, cudaStream_t stream, int splits = 1) {
    constexpr uint32_t stage_bytes = A_TILE_BYTES + BLOCK_N * BLOCK_K * 2;
    constexpr int epi_n = BLOCK_N < 128 ? BLOCK_N : 128;
    const uint32_t staging = BLOCK_M * (epi_n * (p.out_f32 ? 4 : 2) + 16);
    uint32_t pipe = static_cast<uint32_t>(p.num_stages) * stage_bytes;
    if (pipe < staging) pipe = (staging + 1023) / 1024 * 1024;
    p.pipe_bytes = pipe;
    const size_t smem = 1024 + pipe + (2 * MAX_STAGES + 1) * 8 + 16 + BLOCK_N * 4 + (LN ? (2 * BLOCK_N + 4 * BLOCK_M) * 4 : 0);
    // There is deliberately no CTA-pair (tcgen05 cta_group::2) variant of this kernel.  Round 1 had one, and round 2
    // traced the nondeterministic "unspecified launch failure" / hang of round 1 to it: a pair allocates Tensor Memory
    // with tcgen05.alloc.cta_group::2 on both of its SMs, and with ~100 KB of shared memory per CTA another kernel's
    // single-CTA allocator can be resident on ONE of the two SMs -- the pair then never leaves its allocation (flight
    // recorder: exactly the two CTAs of one pair entered and never finished set-up, nothing else in flight; 0 of 7
    // runs hung with the pair GEMM off, 2-4 of 4 with it on, whatever the programmatic-launch settings).  The chain
    // kernels keep their pairs: 226 KB of shared memory and all 512 columns each, so a pair owns both SMs outright,
    // which is also what CUTLASS's 2-SM kernels do.  The single-CTA 128 x 256 tile is faster here anyway (98.2 k vs
    // 92.6 k captions/s with exclusive pairs).
    static cap_device_once smem_once;
    CAP_PROPAGATE(cap_opt_in_smem(smem_once, gemm_tn_bf16_tcgen05<BLOCK_N, STATS, LN>, 200 * 1024));
    CAP_PROPAGATE(install_fault_buffer());
    const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M, tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    if constexpr (LN) {  // the tiles_n CTAs of a row tile form one cluster along x
        cap_launch_kernel(gemm_tn_bf16_tcgen05<BLOCK_N, STATS, LN>, dim3(tiles_n, tiles_m), dim3(GEMM_THREADS), smem,
                          stream, /*cluster_x=*/tiles_n, ta, tb, p);
    } else {
        dim3 grid(tiles_n, tiles_m, splits);
        CAP_LAUNCH((gemm_tn_bf16_tcgen05<BLOCK_N, STATS, LN>), grid, GEMM_THREADS, smem, stream, ta, tb, p);
    }
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("gemm_tn_bf16_tcgen05");
}

unsigned long long* g_gemm_trace = nullptr;

}  // namespace

// tensor-map builder for the other tcgen05 kernels of the library (decode_fused.cu)
namespace cap_gemm {
int make_tmap(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
    return ::make_tmap(map, base, rows, cols, ld, box_rows);
}
}  // namespace cap_gemm

extern "C" int cap_linear(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int out_dtype,
                          int act, int M, int N, int K, cap_stream_t stream) {
    CAP_REQUIRE(x && w && y, "cap_linear: null pointer");
    CAP_REQUIRE(M > 0 && N > 0 && K > 0, "cap_linear: empty problem M=%d N=%d K=%d", M, N, K);
    CAP_REQUIRE(K % 8 == 0 && ldx % 8 == 0, "cap_linear: K=%d and ldx=%d must be multiples of 8 (TMA 16-byte strides)",
                K, ldx);
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "cap_linear: x and w must be 16-byte aligned");
    CAP_REQUIRE(ldx >= K && ldy >= N, "cap_linear: leading dimensions too small");

    const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
    int bn = 32;
    {
        // Widest tile that N fills: with several batches pipelined on separate streams the SMs are kept
        // busy by other kernels, so total L2->smem fill traffic (A is re-read once per N tile) matters
        // more than the CTA count of one GEMM (measured: +8.6 % captions/s vs. "at least one wave").
        // 128x256 tiles pay off when the grid still fills the GPU (encoder, vocabulary: measured 690-820 vs
        // 510-730 TFLOP/s); for the small decode GEMMs 128-wide tiles gave the better end-to-end rate.
        bn = (N >= 256 && tiles_m >= 32) ? 256 : (N >= 128 ? 128 : (N >= 64 ? 64 : 32));
    }

    GemmParams p = {};
    p.out = y;
    p.bias = bias;
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldo = ldy;
    p.out_f32 = (out_dtype == CAP_F32);
    p.act = act;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    // short K loops (<= 8 blocks) get a 2-deep ring: less smem per CTA, more CTAs of concurrent kernels per SM
    int stages = bn == 256 ? 2 : (bn == 128 ? (num_kb <= 8 ? 2 : 3) : 4);
    if (stages > num_kb) stages = num_kb;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 1) stages = 1;
    p.num_stages = stages;
    const size_t esz = p.out_f32 ? 4 : 2;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((static_cast<size_t>(ldy) * esz) % 16 == 0);
    p.trace = g_gemm_trace;

    CUtensorMap ta, tb;
    CAP_PROPAGATE(make_tmap(&ta, x, M, K, ldx, BLOCK_M));
    CAP_PROPAGATE(make_tmap(&tb, w, N, K, K, bn));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (bn) {
        case 256: return launch_gemm<256>(ta, tb, p, s);
        case 128: return launch_gemm<128>(ta, tb, p, s);
        case 64: return launch_gemm<64>(ta, tb, p, s);
        default: return launch_gemm<32>(ta, tb, p, s);
    }
}

// Split-K variant for products with few output tiles and a long contraction (the weight gradients dW = dY^T.X of the
// training step: N x K outputs of 512 .. 2048 over M = thousands of rows): `splits` CTAs per output tile take
// consecutive k-block ranges and write fp32 partial products to partials[z][M][ldy]; the caller sums them
// (cap_sum_partials).  No bias, no activation.
extern "C" int cap_linear_splitk(const void* x, int ldx, const void* w, float* partials, int ldy, int M, int N, int K,
                                 int splits, cap_stream_t stream) {
    CAP_REQUIRE(x && w && partials, "cap_linear_splitk: null pointer");
    CAP_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "cap_linear_splitk: empty problem");
    CAP_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldx >= K && ldy >= N, "cap_linear_splitk: K, ldx multiples of 8; ldx >= K, ldy >= N");
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "cap_linear_splitk: x and w must be 16-byte aligned");
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per = (num_kb + splits - 1) / splits;
    CAP_REQUIRE((splits - 1) * kb_per < num_kb, "cap_linear_splitk: %d splits leave an empty k range (K = %d)", splits, K);
    const int bn = N >= 256 ? 256 : (N >= 128 ? 128 : (N >= 64 ? 64 : 32));
    GemmParams p = {};
    p.out = partials;
    p.bias = nullptr;
    p.M = M; p.N = N; p.K = K; p.ldo = ldy;
    p.out_f32 = 1;
    p.act = CAP_ACT_NONE;
    int stages = bn == 256 ? 2 : (bn == 128 ? 3 : 4);
    if (stages > kb_per) stages = kb_per;
    p.num_stages = stages;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(partials) & 15) == 0) && ((static_cast<size_t>(ldy) * 4) % 16 == 0) &&
               ((static_cast<size_t>(M) * ldy * 4) % 16 == 0);
    p.trace = nullptr;
    p.kb_per_split = kb_per;
    p.split_stride_bytes = static_cast<size_t>(M) * ldy * sizeof(float);
    CUtensorMap ta, tb;
    CAP_PROPAGATE(make_tmap(&ta, x, M, K, ldx, BLOCK_M));
    CAP_PROPAGATE(make_tmap(&tb, w, N, K, K, bn));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (bn) {
        case 256: return launch_gemm<256>(ta, tb, p, s, splits);
        case 128: return launch_gemm<128>(ta, tb, p, s, splits);
        case 64: return launch_gemm<64>(ta, tb, p, s, splits);
        default: return launch_gemm<32>(ta, tb, p, s, splits);
    }
}

// out = LayerNorm(residual + x.w^T + bias) * gamma + beta (+pos), rows in zero_rows zeroed -- one kernel.
extern "C" int cap_linear_layernorm(const void* x, int ldx, const void* w, const float* bias, const float* residual,
                                    int ldr, const float* gamma, const float* beta, float eps, const float* pos,
                                    int pos_rows, const uint8_t* zero_rows, void* out_bf16, int ldo, float* out_f32,
                                    int ldo32, int M, int N, int K, cap_stream_t stream) {
    CAP_REQUIRE(x && w && gamma && beta && (out_bf16 || out_f32), "cap_linear_layernorm: null pointer");
    CAP_REQUIRE(M > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldx >= K, "cap_linear_layernorm: bad shape");
    CAP_REQUIRE(N % 128 == 0 && N >= 128 && N <= 1024, "cap_linear_layernorm: N=%d must be 128..1024 in steps of 128", N);
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "cap_linear_layernorm: x and w must be 16-byte aligned");
    CAP_REQUIRE((!residual || (ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0)) &&
                    (!out_f32 || (ldo32 % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0)) &&
                    (!out_bf16 || (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0)) &&
                    (!pos || (pos_rows > 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0)),
                "cap_linear_layernorm: residual / outputs / pos must be vector-aligned");
    GemmParams p = {};
    p.out = out_bf16;
    p.bias = bias;
    p.M = M; p.N = N; p.K = K; p.ldo = ldo;
    p.out_f32 = 1;  // the staging tile is fp32
    p.act = CAP_ACT_NONE;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    p.num_stages = num_kb <= 8 ? (num_kb < 2 ? num_kb : 2) : 3;
    p.vec_ok = 1;
    p.ln_gamma = gamma; p.ln_beta = beta; p.residual = residual; p.ldr = ldr; p.pos = pos; p.pos_rows = pos_rows;
    p.zero_rows = zero_rows; p.out32 = out_f32; p.ldo32 = ldo32; p.eps = eps;
    p.trace = g_gemm_trace;
    CUtensorMap ta, tb;
    CAP_PROPAGATE(make_tmap(&ta, x, M, K, ldx, BLOCK_M));
    CAP_PROPAGATE(make_tmap(&tb, w, N, K, K, 128));
    return launch_gemm<128, false, true>(ta, tb, p, static_cast<cudaStream_t>(stream));
}

// Vocabulary projection: fp32 logits + per-32-column-chunk log-softmax statistics in one pass.
extern "C" int cap_vocab_logits_stats(const void* x, int ldx, const void* w, const float* bias, float* logits, int ld,
                                      int M, int N, int K, float* part_ms, int* chunks_out, cap_stream_t stream) {
    CAP_REQUIRE(x && w && logits && part_ms, "cap_vocab_logits_stats: null pointer");
    CAP_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldx >= K && ld >= N,
                "cap_vocab_logits_stats: bad shape");
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "cap_vocab_logits_stats: x and w must be 16-byte aligned");
    GemmParams p = {};
    p.out = logits;
    p.bias = bias;
    p.M = M; p.N = N; p.K = K; p.ldo = ld;
    p.out_f32 = 1;
    p.act = CAP_ACT_NONE;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    p.num_stages = num_kb < 2 ? num_kb : 2;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) && (ld % 4 == 0);
    p.part_ms = part_ms;
    p.trace = g_gemm_trace;
    if (chunks_out) *chunks_out = ((N + 255) / 256) * 8;
    CUtensorMap ta, tb;
    CAP_PROPAGATE(make_tmap(&ta, x, M, K, ldx, BLOCK_M));
    CAP_PROPAGATE(make_tmap(&tb, w, N, K, K, 256));
    return launch_gemm<256, true>(ta, tb, p, static_cast<cudaStream_t>(stream));
}

extern "C" int cap_debug_gemm_trace(unsigned long long* device_buffer) {
    g_gemm_trace = device_buffer;
    return CAP_OK;
}

// ------------------------------------------------------------------ CUDA-core cross-check GEMM
namespace {
__global__ void gemm_tn_simt(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                             const float* __restrict__ bias, void* out, int ldo, int out_f32, int act, int M, int N,
                             int K) {
    pdl_prologue();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= N || row >= M) return;
    const bf16* xr = x + static_cast<size_t>(row) * ldx;
    const bf16* wr = w + static_cast<size_t>(col) * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(xr[k]), __bfloat162float(wr[k]), acc);
    if (bias) acc += bias[col];
    acc = apply_act(acc, act);
    if (out_f32)
        reinterpret_cast<float*>(out)[static_cast<size_t>(row) * ldo + col] = acc;
    else
        reinterpret_cast<bf16*>(out)[static_cast<size_t>(row) * ldo + col] = __float2bfloat16_rn(acc);
}
}  // namespace

extern "C" int cap_linear_simt(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy,
                               int out_dtype, int act, int M, int N, int K, cap_stream_t stream) {
    CAP_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, "cap_linear_simt: bad arguments");
    dim3 grid((N + 127) / 128, M);
    CAP_LAUNCH((gemm_tn_simt), grid, 128, 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), ldx, static_cast<const bf16*>(w), bias, y, ldy, out_dtype == CAP_F32, act, M, N, K);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("gemm_tn_simt");
}
