// One decode step of the standard caption decoder as ONE kernel: every CTA owns a tile of 128 beam rows and
// carries it through token embedding, the decoder layers (decoders.py:21-28: self-attention, cross-attention,
// position-wise feed-forward, each followed by residual + LayerNorm) and the vocabulary projection with its
// log-softmax chunk statistics (decoders.py:121-123).  Rows never interact inside a step (beams interact only
// in the selection, beam.cu), so there is no grid-wide dependency: nothing but the weights is shared between
// CTAs, and the ~37 dependent launches of the per-operator path collapse into one.
//
// Structure of a CTA (384 threads; setmaxnreg moves registers from the control warpgroup to the workers):
//   warp 0   : producer -- streams 128x64 weight tiles (TMA, SWIZZLE_128B) through a 4-stage ring, for ALL the
//              step's GEMMs back to back: weights do not depend on activations, so the ring is refilled while
//              the workers are still in an epilogue or an attention phase;
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x128x16, fp32 accumulators in two
//              256-column TMEM buffers: the epilogue of one 256-column chunk overlaps the MMAs of the next);
//   warps 4-11: workers -- TMEM epilogues (bias / ReLU / residual + LayerNorm / log-softmax statistics), the two
//              attention phases on CUDA cores, and the embedding.
// The activation tile is the RESIDENT A operand: 128 rows x 512 columns of bf16 in shared memory as eight
// K-major SWIZZLE_128B k-blocks (the layout TMA would produce), written directly by the epilogues: a thread
// per row puts its 16-byte chunk c of k-block kb at kb*16K + row*128 + ((c ^ (row & 7)) << 4), which is
// bank-conflict-free for a thread-per-row writer and is what the tcgen05.mma descriptor expects.  (A first
// version kept the tile un-swizzled as 8x16-byte core matrices; it was correct but every MMA took ~4x its
// floor -- the tensor pipe was busy fetching A -- see profiles/r01_fused_first_ncu.txt.)
// Only the 2048-wide FFN hidden tile does not fit: it goes through a row-major global scratch buffer and
// comes back through TMA as the streamed A operand of the second FFN GEMM.  The fp32 residual stream lives in
// a global scratch buffer in 16-byte-granule layout [granule][row] (coalesced for a thread-per-row reader).
#include "cap_common.cuh"
#include "tcgen05_ptx.cuh"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

extern std::atomic<long long> g_cap_launches;
namespace cap_gemm {
int make_tmap(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows);
}

namespace {
using namespace cap_ptx;

constexpr int FD = 512;              // d_model = heads * d_k
constexpr int FDFF = 2048;           // feed-forward width
constexpr int FHEADS = 8;
constexpr int TILE_ROWS = 128;
constexpr int NW = 8;                // worker warps
constexpr int FIRST_WORKER_WARP = 4;  // warps 0-3 form the control warpgroup (producer, MMA issuer, two idle)
constexpr int FUSED_THREADS = (FIRST_WORKER_WARP + NW) * 32;
// Worker warps of the chain kernels: 8 or 16 (the epilogues are written for both).  Measured: 16 warps (4 per
// scheduler, 96 registers per thread) are no faster than 8 (84.3 k vs 83.6 k captions/s) -- the chains are paced by
// the weight ring and the shared-memory traffic of the MMAs, not by epilogue instruction latency -- so 8 it is.
constexpr int NW_CHAIN = 8;
constexpr int CHAIN_THREADS = (FIRST_WORKER_WARP + NW_CHAIN) * 32;
constexpr int CHAIN_MAXNREG = NW_CHAIN == 16 ? 96 : 128;
constexpr int STAGE_PITCH_CHAIN = 48;  // 16-warp staging: 32 rows x 32 B (+16 B skew) per warp
constexpr int NB = 4;                // weight ring stages
constexpr uint32_t B_STAGE_BYTES = 128 * 64 * 2;   // one 128-row x 64-column weight tile
constexpr uint32_t A_KB_BYTES = 128 * 64 * 2;      // one k-block of the A operand (8 granules)
constexpr uint32_t GRAN_BYTES = TILE_ROWS * 16;    // one 16-byte granule column of all 128 rows
constexpr int A_SLOTS = 8;           // k-blocks of the resident A tile = slots of the streamed-A ring
constexpr int STAGE_PITCH = 80;      // per-warp store staging: 32 rows x 64 B (+16 B skew)
constexpr int MAXB = 5;              // beams served per image by the cross-attention phase
constexpr int MAX_FUSED_LAYERS = 6;
constexpr int W512_ROWS_PER_LAYER = 3 * FD + FD + FD + FD + FDFF;  // qkv | o1 | q | o2 | w1 (all K = 512)

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + A_SLOTS * A_KB_BYTES;
constexpr uint32_t OFF_STAGE = OFF_B + NB * B_STAGE_BYTES;
constexpr uint32_t STAGE_BYTES_ALL = NW_CHAIN * 32 * STAGE_PITCH_CHAIN > NW * 32 * STAGE_PITCH ? NW_CHAIN * 32 * STAGE_PITCH_CHAIN
                                                                                             : NW * 32 * STAGE_PITCH;
constexpr uint32_t OFF_BIAS = OFF_STAGE + STAGE_BYTES_ALL;
constexpr uint32_t OFF_CBIAS = OFF_BIAS + FD * 4;     // [2][256] chunk biases of the plain projections
constexpr uint32_t OFF_GAMMA = OFF_CBIAS + FD * 4;
constexpr uint32_t OFF_BETA = OFF_GAMMA + FD * 4;
// LayerNorm row statistics [2 (sum, sumsq)][column groups <= 4][128 rows] alias the staging tiles: the LN epilogue
// stages nothing, and a workers_sync separates it from the chunk epilogues on both sides
constexpr uint32_t OFF_STAT = OFF_STAGE;
static_assert(2 * 4 * TILE_ROWS * 4 <= STAGE_BYTES_ALL, "row statistics must fit the staging area");
constexpr uint32_t OFF_BARS = OFF_BETA + FD * 4;
constexpr int NB_PAIR = 8;            // CTA pairs stage HALF of every weight tile: the same 64 KB hold 8 k-block stages
constexpr uint32_t B_STAGE_BYTES_PAIR = 64 * 64 * 2;
constexpr int NUM_BARS = 2 * NB_PAIR + 2 * A_SLOTS + 2 + 2 + 1 + 1 + 1 + 1;
constexpr uint32_t OFF_TMEM = OFF_BARS + NUM_BARS * 8;
constexpr uint32_t FUSED_SMEM = OFF_TMEM + 16;

struct FusedLayerP {
    const float *b_qkv, *b_o1, *g1, *be1, *b_q, *b_o2, *g2, *be2, *b_w1, *b_w2, *g3, *be3;
};

struct FusedParams {
    CUtensorMap map_w512;   // [layers * 5120][512]: per layer qkv | self o | cross q | cross o | ffn1
    CUtensorMap map_w2;     // [layers * 512][2048]
    CUtensorMap map_vocab;  // [V][512]
    CUtensorMap map_h;      // [tiles * 128][2048] FFN hidden scratch
    CUtensorMap map_att;    // [max_rows][512] attention output of the stand-alone attention kernels (chain mode)
    CUtensorMap map_w512_h, map_w2_h, map_vocab_h;   // the three weight maps with 64-row boxes (CTA pairs)
    FusedLayerP layer[MAX_FUSED_LAYERS];
    int n_layers;
    const int32_t* tokens;
    const bf16* word_emb;
    const float* word_pos;
    int pad_idx;
    bf16* qkv_cache;         // [layers][T][R][1536]
    const int32_t* ancestry; // [T][R]
    uint8_t* padflag;        // [T][R]
    const bf16* cross_kv;    // layer l at cross_kv + l * cross_layer_stride: [B][n][K(512) | V(512)]
    size_t cross_layer_stride;
    const uint8_t* enc_mask; // [B][n]
    int n_keys;
    float* res;              // [tiles][128 granules of 4 floats][128 rows][4]   fp32 residual stream
    bf16* qg;                // [tiles][64 granules][128 rows][8]                cross-attention queries
    bf16* hbuf;              // [tiles * 128][2048] row-major                    FFN hidden
    float* logits;
    int ld_logits;
    float* part_ms;          // [R][stat_chunks][2]
    int vocab, vocab_tiles, stat_chunks;
    int t, T, R, B, beam;
    float scale;
    unsigned long long* trace;  // debug: 64 %globaltimer stamps per CTA (cap_debug_fused_trace), else nullptr
    int dbg_skip;            // debug (timing only): bit 0 = issue no MMAs, bit 1 = load no weight tiles
    // chain mode: the kernel runs jobs [job_begin, job_end] of the step's GEMM list only; its first A tile is the
    // embedding (start_embed) or a TMA load of the attention output; the attention itself runs between the chains
    // as stand-alone kernels that share SMs with everything else in flight
    int job_begin, job_end, start_embed;
    bf16* q_out;             // [R][512] cross-attention queries for the stand-alone kernel (chain mode)
    int sparse_logits;       // 1: store only the 32-column groups that can hold one of the row's top-5 candidates
};

__device__ __forceinline__ void fstamp(const FusedParams& p, int tile, int slot, bool who) {
    if (p.trace != nullptr && who) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[static_cast<size_t>(tile) * 64 + slot] = t;
        if (slot == 0 || slot == 1 + 8 * 3 + 2) p.trace[static_cast<size_t>(tile) * 64 + (slot == 0 ? 60 : 61)] = clock64();
    }
}

struct Job {
    const CUtensorMap* map;
    const CUtensorMap* map_half;   // same matrix, 64-row boxes
    int row0, ntiles, kblocks, chunk;
    bool stream;
};

__device__ __forceinline__ Job get_job(const FusedParams& p, int ji) {
    Job j;
    j.kblocks = FD / BLOCK_K;
    j.stream = false;
    j.map = &p.map_w512;
    j.map_half = &p.map_w512_h;
    if (ji == p.n_layers * 6) {
        j.map = &p.map_vocab;
        j.map_half = &p.map_vocab_h;
        j.row0 = 0;
        j.ntiles = p.vocab_tiles;
        j.chunk = 2;
        return j;
    }
    const int L = ji / 6, k = ji % 6;
    const int base = L * W512_ROWS_PER_LAYER;
    switch (k) {
        case 0: j.row0 = base; j.ntiles = 12; j.chunk = 2; break;             // q | k | v of the new token
        case 1: j.row0 = base + 1536; j.ntiles = 4; j.chunk = 4; break;       // self fc_o (+ LN)
        case 2: j.row0 = base + 2048; j.ntiles = 4; j.chunk = 2; break;       // cross fc_q
        case 3: j.row0 = base + 2560; j.ntiles = 4; j.chunk = 4; break;       // cross fc_o (+ LN)
        case 4: j.row0 = base + 3072; j.ntiles = 16; j.chunk = 2; break;      // fc1 (+ ReLU)
        default:
            j.map = &p.map_w2; j.map_half = &p.map_w2_h; j.row0 = L * FD; j.ntiles = 4; j.chunk = 4; j.kblocks = FDFF / BLOCK_K;
            j.stream = true;                                                  // fc2 (+ LN), A = hidden tile
            break;
    }
    return j;
}

// 64 bytes per lane (the lane's row) -> global rows, through the warp's staging tile: every store instruction
// then covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes.
template <bool STREAMING = false>
__device__ __forceinline__ void staged_store64(uint8_t* stage, int lane, const uint4 (&v)[4], uint8_t* gbase,
                                               size_t row_stride, int rows_valid) {
    uint8_t* mine = stage + lane * STAGE_PITCH;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(mine + i * 16) = v[i];
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane;
        const int rr = idx >> 2, part = idx & 3;
        const uint4 x = *reinterpret_cast<const uint4*>(stage + rr * STAGE_PITCH + part * 16);
        if (rr < rows_valid) {
            uint4* dst = reinterpret_cast<uint4*>(gbase + static_cast<size_t>(rr) * row_stride + part * 16);
            if (STREAMING) __stcs(dst, x); else *dst = x;   // logits are written once and almost never read back
        }
    }
    __syncwarp();
}

// byte offset of the 16-byte chunk holding columns [8*chunk, 8*chunk + 8) of `row` inside the resident A tile
__device__ __forceinline__ uint32_t a_tile_off(int row, int chunk) {
    return static_cast<uint32_t>(chunk >> 3) * A_KB_BYTES + static_cast<uint32_t>(row) * 128u +
           (static_cast<uint32_t>((chunk & 7) ^ (row & 7)) << 4);
}

// 16-warp layout: 32 bytes per lane and pass (every store instruction covers 16 rows x 32 contiguous bytes)
template <bool STREAMING = false>
__device__ __forceinline__ void staged_store32(uint8_t* stage, int lane, const uint4& v0, const uint4& v1, uint8_t* gbase,
                                               size_t row_stride, int rows_valid) {
    uint8_t* mine = stage + lane * STAGE_PITCH_CHAIN;
    *reinterpret_cast<uint4*>(mine) = v0;
    *reinterpret_cast<uint4*>(mine + 16) = v1;
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int idx = it * 32 + lane;
        const int rr = idx >> 1, part = idx & 1;
        const uint4 x = *reinterpret_cast<const uint4*>(stage + rr * STAGE_PITCH_CHAIN + part * 16);
        if (rr < rows_valid) {
            uint4* dst = reinterpret_cast<uint4*>(gbase + static_cast<size_t>(rr) * row_stride + part * 16);
            if (STREAMING) __stcs(dst, x); else *dst = x;
        }
    }
    __syncwarp();
}

// 64 bytes per lane to global rows through the warp's staging tile, in the layout of the kernel's warp count
template <bool STREAMING = false>
__device__ __forceinline__ void staged_store(bool wide, uint8_t* stage, int lane, const uint4 (&v)[4], uint8_t* gbase,
                                             size_t row_stride, int rows_valid) {
    if (wide) {
        staged_store64<STREAMING>(stage, lane, v, gbase, row_stride, rows_valid);
    } else {
        staged_store32<STREAMING>(stage, lane, v[0], v[1], gbase, row_stride, rows_valid);
        staged_store32<STREAMING>(stage, lane, v[2], v[3], gbase + 32, row_stride, rows_valid);
    }
}

__device__ __forceinline__ uint4 pack8_u4(const float* f) {
    bf16x8 p = pack8(f);
    return *reinterpret_cast<uint4*>(&p);
}

__device__ __forceinline__ void workers_sync(int nw) { asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory"); }

struct WorkerCtx {
    uint8_t* A_buf;
    uint8_t* stage;   // this warp's staging tile
    float *s_bias, *s_cbias, *s_gamma, *s_beta, *s_stat;
    uint64_t *acc_full, *acc_empty, *a_ready, *h_ready, *ring_free;
    uint8_t* kv_ring;  // this warp's K|V row ring (inside the weight ring)
    bool pair_peer;    // CTA pair, and this CTA is not the leader: consumer-side barriers live in the leader CTA
    uint32_t tmem_base;
    uint32_t use0, use1;  // completed uses of TMEM buffer 0 / 1 (scalars: no dynamically indexed state)
    int toggle;
    int ww, quad, half, lane, wtid;
    int nw, nsub;   // worker warps of this kernel (8 or 16) and column groups per TMEM quadrant (nw / 4); half = ww / 4 in [0, nsub)
    int tile, r0, rows_valid_warp;  // rows of this warp's TMEM quadrant that exist (0..32)
};

__device__ __forceinline__ int acquire_acc(WorkerCtx& c, int chunk) {
    if (chunk == 4) {
        mbar_wait(&c.acc_full[0], c.use0 & 1);
        mbar_wait(&c.acc_full[1], c.use1 & 1);
        tcgen05_fence_after();
        return 0;
    }
    const int b = c.toggle;
    mbar_wait(&c.acc_full[b], (b ? c.use1 : c.use0) & 1);
    tcgen05_fence_after();
    return b;
}

__device__ __forceinline__ void release_acc(WorkerCtx& c, int chunk, int b) {
    tcgen05_fence_before();
    __syncwarp();
    if (chunk == 4) {
        if (c.lane == 0) {
            if (c.pair_peer) { mbar_arrive_remote(&c.acc_empty[0], 0); mbar_arrive_remote(&c.acc_empty[1], 0); }
            else { mbar_arrive(&c.acc_empty[0]); mbar_arrive(&c.acc_empty[1]); }
        }
        c.use0++; c.use1++;
        c.toggle = 0;
    } else {
        if (c.lane == 0) {
            if (c.pair_peer) mbar_arrive_remote(&c.acc_empty[b], 0); else mbar_arrive(&c.acc_empty[b]);
        }
        if (b) c.use1++; else c.use0++;
        c.toggle ^= 1;
    }
}

// the resident A tile was written with ordinary stores: make it visible to tcgen05.mma and tell the issuer
__device__ __forceinline__ void publish_a(WorkerCtx& c) {
    fence_proxy_async_smem();
    __syncwarp();
    if (c.lane == 0) {
        if (c.pair_peer) mbar_arrive_remote(c.a_ready, 0); else mbar_arrive(c.a_ready);
    }
}

enum { EPI_CACHE = 0, EPI_QG = 1, EPI_HID = 2 };

// Epilogue of one 256-column chunk of a plain projection: + bias (ReLU for the hidden layer), bf16, to global.
// `bias_reg` carries this thread's bias value of the chunk across calls: the value of chunk c + 1 is requested
// while chunk c is drained, so the global-load latency is off the per-chunk critical path (n_chunks = chunks of
// the job; the first chunk of a job loads its own value).
// (Measured in round 2 and removed: software-pipelining the four tcgen05.ld of a warp one group ahead of the math
// changed nothing -- 84.5 k vs 85.2 k captions/s -- and storing the bf16 rows straight from registers instead of
// through the staging tile was slower, 81.9 k: the store sectors cost more than the shared-memory round trip.)
template <int KIND>
__device__ __forceinline__ void epilogue_chunk(WorkerCtx& c, const FusedParams& p, const float* bias, int chunk_idx,
                                               int n_chunks, float& bias_reg, bf16* dst_rowmajor, int ld_rowmajor) {
    // stage this chunk's 256 bias values (double-buffered by chunk parity; see the barrier note below)
    const bool tr = (KIND == EPI_HID && chunk_idx == 3 && c.ww == 0 && c.lane == 0);
    fstamp(p, c.tile, 40, tr);
    float* sb = c.s_cbias + (chunk_idx & 1) * 256;
    if (c.wtid < 256) {
        if (chunk_idx == 0) bias_reg = __ldg(bias + c.wtid);
        sb[c.wtid] = bias_reg;
        if (chunk_idx + 1 < n_chunks) bias_reg = __ldg(bias + (chunk_idx + 1) * 256 + c.wtid);
    }
    workers_sync(c.nw);  // a warp reaches the NEXT chunk's barrier only after it has finished reading this one
    fstamp(p, c.tile, 41, tr);
    const int b = acquire_acc(c, 2);
    fstamp(p, c.tile, 42, tr);
    const int row = c.quad * 32 + c.lane;
    const int groups = 8 / c.nsub;   // 32-column groups of the chunk this warp drains (4 with 8 warps, 2 with 16)
    const bool wide = c.nw == NW;
#pragma unroll 1
    for (int i = 0; i < groups; ++i) {
        const int colc = (c.half * groups + i) * 32;
        uint32_t v[32];
        fstamp(p, c.tile, 48, tr && i == 1);
        tmem_ld_32x32b_x32(c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16) + b * 256 + colc, v);
        tmem_ld_wait();
        fstamp(p, c.tile, 49, tr && i == 1);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            f[j] = __uint_as_float(v[j]) + sb[colc + j];
            if (KIND == EPI_HID) f[j] = fmaxf(f[j], 0.f);
        }
        uint4 o[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) o[g] = pack8_u4(f + 8 * g);
        if (tr && i == 1 && o[0].x == 0x12345678u) fstamp(p, c.tile, 63, true);  // keep the math before the stamp
        fstamp(p, c.tile, 50, tr && i == 1);
        const int gcol = chunk_idx * 256 + colc;
        if (KIND == EPI_CACHE) {
            uint8_t* gbase = reinterpret_cast<uint8_t*>(dst_rowmajor + static_cast<size_t>(c.r0 + c.quad * 32) * ld_rowmajor + gcol);
            staged_store(wide, c.stage, c.lane, o, gbase, static_cast<size_t>(ld_rowmajor) * 2, c.rows_valid_warp);
        } else if (KIND == EPI_HID) {
            // row-major scratch [tiles * 128][2048] (whole tiles are allocated: no row guard), re-read by TMA
            uint8_t* gbase = reinterpret_cast<uint8_t*>(p.hbuf + static_cast<size_t>(c.r0 + c.quad * 32) * FDFF + gcol);
            staged_store(wide, c.stage, c.lane, o, gbase, static_cast<size_t>(FDFF) * 2, 32);
        } else {
            // granule layout [tile][granule][row][16 B]: 32 lanes write 512 contiguous bytes per granule
            uint4* base = reinterpret_cast<uint4*>(p.qg) + (static_cast<size_t>(c.tile) * (FD / 8) + gcol / 8) * TILE_ROWS + row;
#pragma unroll
            for (int g = 0; g < 4; ++g) base[static_cast<size_t>(g) * TILE_ROWS] = o[g];
        }
        fstamp(p, c.tile, 43 + i, tr);
    }
    release_acc(c, 2, b);
    fstamp(p, c.tile, 47, tr);
}

// Epilogue of an N = 512 projection followed by residual + LayerNorm (attentions.py:308-309,
// positionwise_feed_forward.py:26): thread = row, the two warps of a TMEM quadrant take 256 columns each and
// exchange (sum, sum of squares); the normalised row goes to the resident A tile (bf16) and back to the fp32
// residual stream (in place: a thread only ever touches its own elements).
// PARK (chain kernels): pass A writes y = projection + bias + residual back into the accumulator's TMEM columns, so
// pass B needs neither the residual (global) nor the bias again; the residual granules of the next 32 columns are
// requested while the current ones are consumed.
template <bool PARK>
__device__ __forceinline__ void epilogue_layernorm(WorkerCtx& c, const FusedParams& p, const float* bias,
                                                   const float* gamma, const float* beta, const uint8_t* zero_rows) {
    for (int i = c.wtid; i < FD; i += c.nw * 32) {
        c.s_bias[i] = __ldg(bias + i);
        c.s_gamma[i] = __ldg(gamma + i);
        c.s_beta[i] = __ldg(beta + i);
    }
    workers_sync(c.nw);
    acquire_acc(c, 4);
    const int row = c.quad * 32 + c.lane;
    const int grow = c.r0 + row;
    const bool live = grow < p.R;
    float4* res = reinterpret_cast<float4*>(p.res) + static_cast<size_t>(c.tile) * (FD / 4) * TILE_ROWS + row;
    const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16);
    float s1 = 0.f, s2 = 0.f;
    if constexpr (PARK) {
    const int groups = 16 / c.nsub;   // 32-column groups of the 512-wide row this warp owns (8 or 4)
#pragma unroll 1
    for (int i = 0; i < groups; ++i) {
        const int c0 = (c.half * groups + i) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        float4 r[8];   // requested while the TMEM load is in flight
#pragma unroll
        for (int g = 0; g < 8; ++g) r[g] = res[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float y0 = __uint_as_float(v[4 * g]) + c.s_bias[c0 + 4 * g] + r[g].x;
            const float y1 = __uint_as_float(v[4 * g + 1]) + c.s_bias[c0 + 4 * g + 1] + r[g].y;
            const float y2 = __uint_as_float(v[4 * g + 2]) + c.s_bias[c0 + 4 * g + 2] + r[g].z;
            const float y3 = __uint_as_float(v[4 * g + 3]) + c.s_bias[c0 + 4 * g + 3] + r[g].w;
            s1 += (y0 + y1) + (y2 + y3);
            s2 = fmaf(y0, y0, s2); s2 = fmaf(y1, y1, s2); s2 = fmaf(y2, y2, s2); s2 = fmaf(y3, y3, s2);
            v[4 * g] = __float_as_uint(y0); v[4 * g + 1] = __float_as_uint(y1);
            v[4 * g + 2] = __float_as_uint(y2); v[4 * g + 3] = __float_as_uint(y3);
        }
        tmem_st_32x32b_x32(taddr + c0, v);
    }
    tmem_st_wait();
    c.s_stat[c.half * TILE_ROWS + row] = s1;
    c.s_stat[(4 + c.half) * TILE_ROWS + row] = s2;
    asm volatile("bar.sync %0, %1;" ::"r"(2 + c.quad), "r"(c.nsub * 32) : "memory");  // the quadrant's column groups
    float S1 = 0.f, S2 = 0.f;
    for (int k = 0; k < c.nsub; ++k) {
        S1 += c.s_stat[k * TILE_ROWS + row];
        S2 += c.s_stat[(4 + k) * TILE_ROWS + row];
    }
    const float mean = S1 * (1.f / FD);
    const float var = fmaxf(S2 * (1.f / FD) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const bool zero = !live || (zero_rows != nullptr && zero_rows[grow] != 0);
#pragma unroll 1
    for (int i = 0; i < groups; ++i) {
        const int c0 = (c.half * groups + i) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
            f[j] = zero ? 0.f : (__uint_as_float(v[j]) - mean) * rstd * c.s_gamma[c0 + j] + c.s_beta[c0 + j];
#pragma unroll
        for (int g = 0; g < 8; ++g)
            res[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row, c0 / 8 + g)) = pack8_u4(f + 8 * g);
    }
    } else {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int c0 = c.half * 256 + i * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        float4 r4[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) r4[g] = res[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float y0 = __uint_as_float(v[4 * g]) + c.s_bias[c0 + 4 * g] + r4[g].x;
            const float y1 = __uint_as_float(v[4 * g + 1]) + c.s_bias[c0 + 4 * g + 1] + r4[g].y;
            const float y2 = __uint_as_float(v[4 * g + 2]) + c.s_bias[c0 + 4 * g + 2] + r4[g].z;
            const float y3 = __uint_as_float(v[4 * g + 3]) + c.s_bias[c0 + 4 * g + 3] + r4[g].w;
            s1 += (y0 + y1) + (y2 + y3);
            s2 = fmaf(y0, y0, s2); s2 = fmaf(y1, y1, s2); s2 = fmaf(y2, y2, s2); s2 = fmaf(y3, y3, s2);
        }
    }
    c.s_stat[c.half * TILE_ROWS + row] = s1;
    c.s_stat[(2 + c.half) * TILE_ROWS + row] = s2;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + c.quad) : "memory");
    const float S1 = c.s_stat[row] + c.s_stat[TILE_ROWS + row];
    const float S2 = c.s_stat[2 * TILE_ROWS + row] + c.s_stat[3 * TILE_ROWS + row];
    const float mean = S1 * (1.f / FD);
    const float var = fmaxf(S2 * (1.f / FD) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const bool zero = !live || (zero_rows != nullptr && zero_rows[grow] != 0);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int c0 = c.half * 256 + i * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        float4 r4[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) r4[g] = res[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float rr[4] = {r4[g].x, r4[g].y, r4[g].z, r4[g].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const float y = __uint_as_float(v[j]) + c.s_bias[c0 + j] + rr[e];
                f[j] = zero ? 0.f : (y - mean) * rstd * c.s_gamma[c0 + j] + c.s_beta[c0 + j];
            }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
            res[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row, c0 / 8 + g)) = pack8_u4(f + 8 * g);
    }
    }
    release_acc(c, 4, 0);
    publish_a(c);
}

// x = Emb[token] + pos[t + 1] (decoders.py:107-112); pad flag of the fed token; warp per row.
__device__ __forceinline__ void embed_phase(WorkerCtx& c, const FusedParams& p) {
    uint8_t* pad_t = p.padflag + static_cast<size_t>(p.t) * p.R;
    const float* pos = p.word_pos + static_cast<size_t>(p.t + 1) * FD + c.lane * 16;
    float pv[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(pos) + i);
        pv[4 * i] = t4.x; pv[4 * i + 1] = t4.y; pv[4 * i + 2] = t4.z; pv[4 * i + 3] = t4.w;
    }
    float4* res = reinterpret_cast<float4*>(p.res) + static_cast<size_t>(c.tile) * (FD / 4) * TILE_ROWS;
    const int ROWS_PER_WARP = TILE_ROWS / c.nw;
    // lane i fetches the token of the warp's row i; rows are then embedded a few at a time (many loads in flight)
    int mytok = -1;
    if (c.lane < ROWS_PER_WARP) {
        const int grow = c.r0 + c.ww * ROWS_PER_WARP + c.lane;
        if (grow < p.R) {
            mytok = p.tokens[grow];
            pad_t[grow] = (mytok == p.pad_idx) ? 1 : 0;
        }
    }
    constexpr int EMB_ROWS = 4;   // rows embedded per round: 8 independent 16-byte loads in flight per lane
#pragma unroll 1
    for (int i0 = 0; i0 < ROWS_PER_WARP; i0 += EMB_ROWS) {
        bf16x8 e[EMB_ROWS][2];
        int tk[EMB_ROWS];
#pragma unroll
        for (int u = 0; u < EMB_ROWS; ++u) {
            tk[u] = __shfl_sync(0xffffffffu, mytok, i0 + u);
            const bf16x8* ep = reinterpret_cast<const bf16x8*>(p.word_emb + static_cast<size_t>(max(tk[u], 0)) * FD + c.lane * 16);
            e[u][0] = ep[0];
            e[u][1] = ep[1];
        }
#pragma unroll
        for (int u = 0; u < EMB_ROWS; ++u) {
            const int row = c.ww * ROWS_PER_WARP + i0 + u;
            float a[16];
            unpack8(e[u][0], a);
            unpack8(e[u][1], a + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = tk[u] >= 0 ? a[j] + pv[j] : 0.f;
#pragma unroll
            for (int g = 0; g < 2; ++g)
                *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row, 2 * c.lane + g)) = pack8_u4(a + 8 * g);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                res[static_cast<size_t>(4 * c.lane + g) * TILE_ROWS + row] = make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
        }
    }
}

// ------------------------------------------------------------------------------ streamed K|V rows
// Both attention phases read 2 KB rows (K | V of one key, contiguous in HBM) that nothing else on the SM
// needs.  Each worker warp streams them with cp.async through a private 4-slot ring carved out of the weight
// ring (idle during an attention phase: the producer is gated until the phase ends), three rows in flight
// while the fourth is consumed -- the loads are decoupled from the arithmetic and cost no registers.
// Chunk k (16 bytes) of a 1 KB half-row is stored at position k/2 + 32*(k%2): lane l then reads its 16
// columns as two conflict-free 16-byte accesses (positions l and 32 + l).
constexpr int KV_SLOTS = 4;
constexpr uint32_t KV_ROW_BYTES = 2 * FD * 2;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc))
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void kv_issue(uint8_t* slot, const bf16* src, int lane) {
    const uint8_t* g = reinterpret_cast<const uint8_t*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = lane + 32 * i;
        const int kk = k & 63;
        const int pos = (k & 64) + (kk >> 1) + ((kk & 1) << 5);
        cp_async_16(slot + pos * 16, g + k * 16);
    }
}

__device__ __forceinline__ void kv_read(const uint8_t* slot, int lane, float* kf, float* vf) {
    const bf16x8* s8 = reinterpret_cast<const bf16x8*>(slot);
    unpack8(s8[lane], kf);
    unpack8(s8[32 + lane], kf + 8);
    unpack8(s8[64 + lane], vf);
    unpack8(s8[96 + lane], vf + 8);
}

// Stateful self-attention of the new token over its beam history (attentions.py:297-304 with the running
// mask of decoders.py:101-103): warp per row, lanes tile the 512-wide row (4 lanes per head), keys found
// through the ancestry table, online softmax; output straight into the resident A tile.  The (row, key)
// pairs of the warp's 16 rows form ONE stream of K|V rows, so the pipeline never drains between rows.
__device__ __forceinline__ void self_attention_phase(WorkerCtx& c, const FusedParams& p, const bf16* cache_l) {
    constexpr int EPL = 16;
    constexpr int ROWS_PER_WARP = TILE_ROWS / NW;
    const int t = p.t, R = p.R;
    const size_t row_stride = 3 * FD;
    const size_t step_stride = static_cast<size_t>(R) * row_stride;
    const int nkeys = t + 1;
    const int row0 = c.ww * ROWS_PER_WARP;
    const int nrows = max(0, min(ROWS_PER_WARP, R - (c.r0 + row0)));
    const int items = nrows * nkeys;
    // slot | pad << 31 of every (row, key) of this warp, gathered with all lanes in parallel
    int32_t* meta = reinterpret_cast<int32_t*>(c.stage);
    for (int idx = c.lane; idx < items; idx += 32) {
        const int i = idx / nkeys, j = idx - i * nkeys;
        const int r = c.r0 + row0 + i;
        const int sl = (j == t) ? r : p.ancestry[static_cast<size_t>(j) * R + r];
        const int pad = p.padflag[static_cast<size_t>(j) * R + sl] != 0;
        meta[idx] = sl | (pad << 31);
    }
    __syncwarp();
    uint8_t* ring = c.kv_ring;
    auto src_of = [&](int idx) -> const bf16* {
        const int i = idx / nkeys, j = idx - i * nkeys;
        return cache_l + j * step_stride + static_cast<size_t>(meta[idx] & 0x7fffffff) * row_stride + FD;
    };
#pragma unroll
    for (int n = 0; n < KV_SLOTS - 1; ++n) {
        if (n < items) kv_issue(ring + n * KV_ROW_BYTES, src_of(n), c.lane);
        cp_async_commit();
    }
    float q[EPL], acc[EPL], m = -INFINITY, l = 0.f;
    int i = 0, j = 0;
#pragma unroll 1
    for (int n = 0; n < items; ++n) {
        if (n + KV_SLOTS - 1 < items) kv_issue(ring + ((n + KV_SLOTS - 1) % KV_SLOTS) * KV_ROW_BYTES, src_of(n + KV_SLOTS - 1), c.lane);
        cp_async_commit();
        if (j == 0) {
            const bf16x8* qp = reinterpret_cast<const bf16x8*>(cache_l + t * step_stride + static_cast<size_t>(c.r0 + row0 + i) * row_stride + c.lane * EPL);
            unpack8(qp[0], q);
            unpack8(qp[1], q + 8);
#pragma unroll
            for (int e = 0; e < EPL; ++e) { q[e] *= p.scale; acc[e] = 0.f; }
            m = -INFINITY;
            l = 0.f;
        }
        cp_async_wait<KV_SLOTS - 1>();
        __syncwarp();  // every lane's copies of item n have landed
        if (meta[n] >= 0) {  // key not fed <pad> (warp-uniform)
            float kf[EPL], vf[EPL];
            kv_read(ring + (n % KV_SLOTS) * KV_ROW_BYTES, c.lane, kf, vf);
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) s = fmaf(q[e], kf[e], s);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            const float m_new = fmaxf(m, s);
            const float corr = __expf(m - m_new);
            const float pr = __expf(s - m_new);
            l = l * corr + pr;
#pragma unroll
            for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * corr + pr * vf[e];
            m = m_new;
        }
        __syncwarp();  // slot n % KV_SLOTS may be refilled by the next iteration's issue
        if (++j == nkeys) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) acc[e] *= inv;
#pragma unroll
            for (int g = 0; g < 2; ++g)
                *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row0 + i, 2 * c.lane + g)) = pack8_u4(acc + 8 * g);
            j = 0;
            ++i;
        }
    }
    cp_async_wait<0>();
}

// Cross-attention over the image's cached K|V (decoders.py:23): warp per image, all heads and all of the
// image's beams at once (lane owns 16 consecutive columns, 4 lanes per head), so every K/V row crosses HBM
// once per step; the (image, key) pairs of the warp's images form one stream of K|V rows.
__device__ __forceinline__ void cross_attention_phase(WorkerCtx& c, const FusedParams& p, const bf16* kv_l) {
    constexpr int EPL = 16;
    const int beam = p.beam, n = p.n_keys;
    const int img_lo = c.r0 / beam;
    const int last_row = min(c.r0 + TILE_ROWS, p.R) - 1;
    const int img_hi = last_row / beam;
    const int first = img_lo + c.ww;
    const int nimg = first <= img_hi ? (img_hi - first) / NW + 1 : 0;
    const int items = nimg * n;
    const uint4* qg = reinterpret_cast<const uint4*>(p.qg) + static_cast<size_t>(c.tile) * (FD / 8) * TILE_ROWS;
    uint8_t* ring = c.kv_ring;
    auto src_of = [&](int idx) -> const bf16* {
        const int ii = idx / n, j = idx - ii * n;
        return kv_l + (static_cast<size_t>(first + ii * NW) * n + j) * 2 * FD;
    };
#pragma unroll
    for (int k = 0; k < KV_SLOTS - 1; ++k) {
        if (k < items) kv_issue(ring + k * KV_ROW_BYTES, src_of(k), c.lane);
        cp_async_commit();
    }
    bf16x8 qreg[MAXB][2];
    float m[MAXB], l[MAXB], acc[MAXB][EPL];
    uint32_t maskbits = 0;  // bit i: key lane + 32*i of the current image is padding
    int ii = 0, j = 0, nb = 0, lr0 = 0;
#pragma unroll 1
    for (int k = 0; k < items; ++k) {
        if (k + KV_SLOTS - 1 < items) kv_issue(ring + ((k + KV_SLOTS - 1) % KV_SLOTS) * KV_ROW_BYTES, src_of(k + KV_SLOTS - 1), c.lane);
        cp_async_commit();
        if (j == 0) {
            const int img = first + ii * NW;
            const int row_begin = max(img * beam, c.r0);
            nb = min(img * beam + beam, last_row + 1) - row_begin;
            lr0 = row_begin - c.r0;
#pragma unroll
            for (int bb = 0; bb < MAXB; ++bb) {
                const int lr = lr0 + min(bb, nb - 1);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint4 u = qg[static_cast<size_t>(2 * c.lane + g) * TILE_ROWS + lr];
                    qreg[bb][g] = *reinterpret_cast<const bf16x8*>(&u);
                }
                m[bb] = -INFINITY;
                l[bb] = 0.f;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[bb][e] = 0.f;
            }
            maskbits = 0;
            if (p.enc_mask != nullptr) {
                const uint8_t* mrow = p.enc_mask + static_cast<size_t>(img) * n;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int key = c.lane + 32 * w;
                    if (key < n && mrow[key]) maskbits |= 1u << w;
                }
            }
        }
        const bool masked = ((__shfl_sync(0xffffffffu, maskbits, j & 31) >> (j >> 5)) & 1u) != 0;
        cp_async_wait<KV_SLOTS - 1>();
        __syncwarp();
        if (!masked) {  // warp-uniform
            float kf[EPL], vf[EPL];
            kv_read(ring + (k % KV_SLOTS) * KV_ROW_BYTES, c.lane, kf, vf);
#pragma unroll
            for (int bb = 0; bb < MAXB; ++bb) {
                if (bb >= nb) continue;  // warp-uniform
                float qf[EPL];
                unpack8(qreg[bb][0], qf);
                unpack8(qreg[bb][1], qf + 8);
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < EPL; ++e) s = fmaf(qf[e], kf[e], s);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s *= p.scale;
                const float m_new = fmaxf(m[bb], s);
                const float corr = __expf(m[bb] - m_new);
                const float pr = __expf(s - m_new);
                l[bb] = l[bb] * corr + pr;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[bb][e] = acc[bb][e] * corr + pr * vf[e];
                m[bb] = m_new;
            }
        }
        __syncwarp();
        if (++j == n) {
#pragma unroll
            for (int bb = 0; bb < MAXB; ++bb) {
                if (bb >= nb) continue;
                const float inv = l[bb] > 0.f ? 1.f / l[bb] : 0.f;
                float o[EPL];
#pragma unroll
                for (int e = 0; e < EPL; ++e) o[e] = acc[bb][e] * inv;
#pragma unroll
                for (int g = 0; g < 2; ++g)
                    *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(lr0 + bb, 2 * c.lane + g)) = pack8_u4(o + 8 * g);
            }
            j = 0;
            ++ii;
        }
    }
    cp_async_wait<0>();
}

// end of an attention phase: the resident A tile is complete and the weight ring is handed back
__device__ __forceinline__ void publish_attention(WorkerCtx& c) {
    fence_proxy_async_smem();
    __syncwarp();
    if (c.lane == 0) {
        mbar_arrive(c.a_ready);
        mbar_arrive(c.ring_free);
    }
}

// Vocabulary projection epilogue (bias-free fc, decoders.py:90,121): fp32 logits + per-32-column-chunk
// (max, sum exp(x - max)) for beam_chunkmerge_kernel (beam.cu).
// Sparse mode: the selection (beam_chunkmerge_kernel) reads back only the `beam` <= 5 groups of 32 columns with
// the largest maxima of a row.  A thread (= row, one half of every chunk) keeps the five largest group maxima it
// has seen; a group is stored only if its maximum reaches the fifth of them -- every group of the row's final
// top five passes that test when it is produced, and ~85 % of the 52 MB of logits per step are never written.
// (the per-group body is vocab_group below)
__device__ __forceinline__ void vocab_group(WorkerCtx& c, const FusedParams& p, int chunk_idx, int colc, int grow, bool wide,
                                            const uint32_t (&v)[32], float (&top)[5]);

__device__ __forceinline__ void epilogue_vocab_chunk(WorkerCtx& c, const FusedParams& p, int chunk_idx, float (&top)[5]) {
    const int b = acquire_acc(c, 2);
    const int row = c.quad * 32 + c.lane;
    const int grow = c.r0 + row;
    const int groups = 8 / c.nsub;   // 32-column groups of the chunk this warp drains
    const bool wide = c.nw == NW;
#pragma unroll 1
    for (int i = 0; i < groups; ++i) {
        const int colc = (c.half * groups + i) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16) + b * 256 + colc, v);
        tmem_ld_wait();
        vocab_group(c, p, chunk_idx, colc, grow, wide, v, top);
    }
    release_acc(c, 2, b);
}

// statistics + (sparse) store of one 32-column group of the vocabulary projection
__device__ __forceinline__ void vocab_group(WorkerCtx& c, const FusedParams& p, int chunk_idx, int colc, int grow, bool wide,
                                            const uint32_t (&v)[32], float (&top)[5]) {
    {
        const int gcol = chunk_idx * 256 + colc;
        const int valid = p.vocab - gcol;  // columns of this 32-chunk inside the vocabulary (warp-uniform)
        float cm = -INFINITY, cs = 0.f;
        if (valid >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
            const float cm2 = cm * 1.4426950408889634f;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += exp2f(fmaf(__uint_as_float(v[j]), 1.4426950408889634f, -cm2));
        } else if (valid > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = fmaxf(cm, j < valid ? __uint_as_float(v[j]) : -INFINITY);
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += j < valid ? __expf(__uint_as_float(v[j]) - cm) : 0.f;
        }
        if (grow < p.R)
            *reinterpret_cast<float2*>(p.part_ms + (static_cast<size_t>(grow) * p.stat_chunks + chunk_idx * 8 + colc / 32) * 2) =
                make_float2(cm, cs);
        if (gcol < p.ld_logits) {
            if (p.sparse_logits) {
                const bool keep = cm >= top[4] && grow < p.R;
                float x = cm;   // insert into the descending top five
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float hi = fmaxf(top[k], x);
                    x = fminf(top[k], x);
                    top[k] = hi;
                }
                if (keep) {
                    uint4* dst = reinterpret_cast<uint4*>(p.logits + static_cast<size_t>(grow) * p.ld_logits + gcol);
#pragma unroll
                    for (int g = 0; g < 8; ++g) __stcs(dst + g, make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
                }
            } else {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint4 o[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        o[g] = make_uint4(v[hh * 16 + 4 * g], v[hh * 16 + 4 * g + 1], v[hh * 16 + 4 * g + 2], v[hh * 16 + 4 * g + 3]);
                    uint8_t* gbase = reinterpret_cast<uint8_t*>(p.logits + static_cast<size_t>(c.r0 + c.quad * 32) * p.ld_logits + gcol + hh * 16);
                    staged_store<true>(wide, c.stage, c.lane, o, gbase, static_cast<size_t>(p.ld_logits) * 4, c.rows_valid_warp);
                }
            }
        }
    }
}

// CHAIN = false: the whole step (attention phases on this CTA's CUDA cores).  CHAIN = true: a range of the
// step's GEMM jobs with their epilogues (see FusedParams::job_begin).
// The chain instantiation is capped at 128 registers per thread (setmaxnreg then moves them: 64 for the control
// warps, 160 for the epilogue warps): a chain CTA then takes 3/4 of the SM's register file instead of all of it, so
// it can start on an SM that still hosts a few small CTAs of other streams' kernels, and vice versa.
// PAIR (chains only): two CTAs of a cluster, each with its own 128-row tile, share every weight tile -- each
// stages HALF of it (64 of the 128 rows), the leader issues tcgen05.mma.cta_group::2 (M = 256) over both shared
// memories and both CTAs drain their own accumulator rows.  A ring byte then feeds 256 rows instead of 128: the
// same 64 KB ring keeps 8 k-block stages in flight instead of 4 (the chains are bound by that ring's latency).
template <bool CHAIN, bool PAIR = false>
__global__ void __launch_bounds__(CHAIN ? CHAIN_THREADS : FUSED_THREADS, 1) __maxnreg__(CHAIN ? CHAIN_MAXNREG : 168)
decode_step_fused_kernel(const __grid_constant__ FusedParams p) {
    constexpr int NWK = CHAIN ? NW_CHAIN : NW;   // worker warps of this instantiation
    static_assert(CHAIN || !PAIR, "CTA pairs exist for the chain kernels only");
    constexpr int NBX = PAIR ? NB_PAIR : NB;
    constexpr uint32_t BST = PAIR ? B_STAGE_BYTES_PAIR : B_STAGE_BYTES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    // 1024-byte alignment (SWIZZLE_128B tiles) comes from the declaration: rounding the address up by hand
    // goes through an integer and makes every later access a generic LD/ST instead of LDS/STS.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* A_buf = smem + OFF_A;
    uint8_t* B_ring = smem + OFF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* b_full = bars;
    uint64_t* b_empty = b_full + NB_PAIR;
    uint64_t* a_full = b_empty + NB_PAIR;
    uint64_t* a_empty = a_full + A_SLOTS;
    uint64_t* acc_full = a_empty + A_SLOTS;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* a_ready = acc_empty + 2;
    uint64_t* h_ready = a_ready + 1;
    uint64_t* ring_free = h_ready + 1;
    uint64_t* a_load = ring_free + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int job_lo = CHAIN ? p.job_begin : 0;
    const int job_hi = CHAIN ? p.job_end : p.n_layers * 6;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_w512)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_w2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_vocab)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_h)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            constexpr int consumers = PAIR ? 2 * NWK : NWK;   // pair: the leader's barriers hear both CTAs' workers
            for (int s = 0; s < NBX; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            for (int s = 0; s < A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], consumers); }
            mbar_init(a_ready, consumers);
            mbar_init(h_ready, NWK);
            mbar_init(ring_free, NWK);
            mbar_init(a_load, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (PAIR) tmem_alloc_2sm<512>(tmem_slot); else tmem_alloc<512>(tmem_slot);
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync(); else __syncthreads();   // pair: the peer's barriers exist before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < FIRST_WORKER_WARP) {
    // control warpgroup: hand registers to the workers (the role split must sit INSIDE this branch so that
    // the register limit of each region is unambiguous to ptxas)
    if constexpr (CHAIN) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");   // 128 x 40 + 512 x 104 regs
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");                    // to squeeze these warps
    if (warp == 0) {
        // ------------------------------------------------------------------ producer (weights never wait)
        // The whole warp runs the loop (uniform control flow); one elected lane issues the copies.
        uint32_t bcount = 0, acount = 0, hphase = 0, rphase = 0;
        const uint64_t keep_policy = l2_policy_evict_last();   // weights: shared by every tile of every batch in flight
        for (int ji = job_lo; ji <= job_hi; ++ji) {
            const Job job = get_job(p, ji);
            const int nch = job.ntiles / job.chunk;
            if (!CHAIN && ji < p.n_layers * 6 && (ji % 6 == 1 || ji % 6 == 3)) {
                // fc_o follows an attention phase, which borrows the weight ring for its K|V rows
                mbar_wait(ring_free, rphase);
                rphase ^= 1;
            }
            for (int c = 0; c < nch; ++c) {
                for (int kb = 0; kb < job.kblocks; ++kb) {
                    for (int j = 0; j < job.chunk; ++j) {
                        const uint32_t s = bcount % NBX;
                        mbar_wait(&b_empty[s], ((bcount / NBX) & 1) ^ 1);
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                // both CTAs signal the LEADER's barrier, which expects the two halves of the tile
                                if (leader) mbar_arrive_expect_tx(&b_full[s], 2 * BST);
                                tma_load_2d_2sm_hint(B_ring + s * BST, job.map_half, &b_full[s], kb * BLOCK_K,
                                                     job.row0 + (c * job.chunk + j) * 128 + static_cast<int>(rank) * 64,
                                                     keep_policy);
                            } else if (p.dbg_skip & 2) {
                                mbar_arrive(&b_full[s]);
                            } else {
                                mbar_arrive_expect_tx(&b_full[s], B_STAGE_BYTES);
                                tma_load_2d_hint(B_ring + s * B_STAGE_BYTES, job.map, &b_full[s], kb * BLOCK_K,
                                                 job.row0 + (c * job.chunk + j) * 128, keep_policy);
                            }
                        }
                        __syncwarp();
                        ++bcount;
                    }
                    if (job.stream) {
                        if (kb == 0) {  // the workers have written (and proxy-fenced) the whole hidden tile
                            mbar_wait(h_ready, hphase);
                            hphase ^= 1;
                        }
                        const uint32_t slot = acount % A_SLOTS;
                        mbar_wait(&a_empty[slot], ((acount / A_SLOTS) & 1) ^ 1);
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                if (leader) mbar_arrive_expect_tx(&a_full[slot], 2 * A_KB_BYTES);
                                tma_load_2d_2sm(A_buf + slot * A_KB_BYTES, &p.map_h, &a_full[slot], kb * BLOCK_K, tile * TILE_ROWS);
                            } else {
                                mbar_arrive_expect_tx(&a_full[slot], A_KB_BYTES);
                                tma_load_2d(A_buf + slot * A_KB_BYTES, &p.map_h, &a_full[slot], kb * BLOCK_K, tile * TILE_ROWS);
                            }
                        }
                        __syncwarp();
                        ++acount;
                    }
                }
            }
        }
        pdl_launch_dependents();
    } else if (warp == 1 && leader) {
        // ------------------------------------------------------------------ MMA issuer (pair: the leader's only)
        // Uniform control flow for the whole warp; the elected lane issues tcgen05.mma / tcgen05.commit.
        constexpr uint32_t idesc = make_instr_desc(PAIR ? 256 : 128, 128);
        auto wait_consumers_at = [](uint64_t* bar, uint32_t parity, int line) {
            if constexpr (PAIR) mbar_wait_cluster_impl(bar, parity, line); else mbar_wait_impl(bar, parity, line);
        };
#define wait_consumers(bar, parity) wait_consumers_at(bar, parity, __LINE__)
        auto commit = [](uint64_t* bar) {
            if constexpr (PAIR) umma_commit_2sm(bar); else umma_commit(bar);
        };
        uint32_t bcount = 0, acount = 0, use0 = 0, use1 = 0, ar = 0;
        int toggle = 0;
        for (int ji = job_lo; ji <= job_hi; ++ji) {
            const Job job = get_job(p, ji);
            const int nch = job.ntiles / job.chunk;
            if (CHAIN && ji == job_lo && !p.start_embed) {
                mbar_wait(a_load, 0);   // the chain's first A tile arrives by TMA (warp 2)
            } else if (!job.stream) {
                wait_consumers(a_ready, ar & 1);
                ++ar;
            }
            tcgen05_fence_after();
            for (int c = 0; c < nch; ++c) {
                int b = 0;
                uint32_t colbase = 0;
                if (job.chunk == 4) {
                    wait_consumers(&acc_empty[0], (use0 & 1) ^ 1);
                    wait_consumers(&acc_empty[1], (use1 & 1) ^ 1);
                } else {
                    b = toggle;
                    wait_consumers(&acc_empty[b], ((b ? use1 : use0) & 1) ^ 1);
                    colbase = b * 256;
                }
                tcgen05_fence_after();
                for (int kb = 0; kb < job.kblocks; ++kb) {
                    const uint8_t* a_tile;
                    uint32_t slot = 0;
                    if (job.stream) {
                        slot = acount % A_SLOTS;
                        mbar_wait(&a_full[slot], (acount / A_SLOTS) & 1);
                        a_tile = A_buf + slot * A_KB_BYTES;
                    } else {
                        a_tile = A_buf + kb * A_KB_BYTES;
                    }
                    const uint64_t a_desc = make_smem_desc(a_tile);
                    for (int j = 0; j < job.chunk; ++j) {
                        const uint32_t s = bcount % NBX;
                        mbar_wait(&b_full[s], (bcount / NBX) & 1);
                        tcgen05_fence_after();
                        const uint64_t b_desc = make_smem_desc(B_ring + s * BST);
                        if (elect_one_sync()) {
                            if (!(p.dbg_skip & 1)) {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                    if constexpr (PAIR)
                                        umma_bf16_2sm(tmem_base + colbase + j * 128, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                                      (kb | k) != 0 ? 1u : 0u);
                                    else
                                        umma_bf16(tmem_base + colbase + j * 128, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                                  (kb | k) != 0 ? 1u : 0u);
                                }
                            }
                            commit(&b_empty[s]);  // stage reusable (in both CTAs) once these MMAs have read it
                        }
                        __syncwarp();
                        ++bcount;
                    }
                    if (job.stream) {
                        if (elect_one_sync()) commit(&a_empty[slot]);
                        __syncwarp();
                        ++acount;
                    }
                }
                if (elect_one_sync()) {
                    if (job.chunk == 4) {
                        commit(&acc_full[0]);
                        commit(&acc_full[1]);
                    } else {
                        commit(&acc_full[b]);
                    }
                }
                __syncwarp();
                if (job.chunk == 4) {
                    ++use0; ++use1;
                    toggle = 0;
                } else {
                    if (b) ++use1; else ++use0;
                    toggle ^= 1;
                }
            }
        }
#undef wait_consumers
        pdl_launch_dependents();
    } else if (CHAIN && warp == 2) {
        // ------------------------------------------------------------------ A-tile loader (chain mode)
        if (!p.start_embed) {
            pdl_wait();  // the attention kernel that wrote the tile is the stream predecessor
            if (elect_one_sync()) {
                if constexpr (PAIR) {
                    if (leader) mbar_arrive_expect_tx(a_load, 2 * A_SLOTS * A_KB_BYTES);
                    for (int kb = 0; kb < A_SLOTS; ++kb)
                        tma_load_2d_2sm(A_buf + kb * A_KB_BYTES, &p.map_att, a_load, kb * BLOCK_K, tile * TILE_ROWS);
                } else {
                    mbar_arrive_expect_tx(a_load, A_SLOTS * A_KB_BYTES);
                    for (int kb = 0; kb < A_SLOTS; ++kb)
                        tma_load_2d(A_buf + kb * A_KB_BYTES, &p.map_att, a_load, kb * BLOCK_K, tile * TILE_ROWS);
                }
            }
            __syncwarp();
        }
    }
    } else {
        // ------------------------------------------------------------------ workers
        if constexpr (CHAIN && NW_CHAIN == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
        else if constexpr (CHAIN) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        WorkerCtx c;
        c.A_buf = A_buf;
        c.ww = warp - FIRST_WORKER_WARP;
        c.quad = warp & 3;
        c.half = c.ww >> 2;
        c.lane = lane;
        c.wtid = threadIdx.x - FIRST_WORKER_WARP * 32;
        c.nw = NWK;
        c.nsub = NWK / 4;
        c.stage = smem + OFF_STAGE + c.ww * 32 * (NWK == NW ? STAGE_PITCH : STAGE_PITCH_CHAIN);   // matches staged_store's `wide`
        c.s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
        c.s_cbias = reinterpret_cast<float*>(smem + OFF_CBIAS);
        c.s_gamma = reinterpret_cast<float*>(smem + OFF_GAMMA);
        c.s_beta = reinterpret_cast<float*>(smem + OFF_BETA);
        c.s_stat = reinterpret_cast<float*>(smem + OFF_STAT);
        c.acc_full = acc_full;
        c.acc_empty = acc_empty;
        c.a_ready = a_ready;
        c.h_ready = h_ready;
        c.ring_free = ring_free;
        c.kv_ring = B_ring + c.ww * KV_SLOTS * KV_ROW_BYTES;
        c.pair_peer = PAIR && !leader;
        c.tmem_base = tmem_base;
        c.use0 = c.use1 = 0;
        c.toggle = 0;
        c.tile = tile;
        c.r0 = tile * TILE_ROWS;
        c.rows_valid_warp = max(0, min(32, p.R - (c.r0 + c.quad * 32)));

        float bias_reg = 0.f;
        if constexpr (CHAIN) {
            pdl_wait();  // everything this chain reads was written by stream predecessors
            uint8_t* pad_t = p.padflag + static_cast<size_t>(p.t) * p.R;
            if (p.start_embed) {
                embed_phase(c, p);
                workers_sync(c.nw);
                publish_a(c);
            }
            for (int ji = job_lo; ji <= job_hi; ++ji) {
                if (ji == p.n_layers * 6) {
                    pdl_launch_dependents();
                    const int vchunks = p.vocab_tiles / 2;
                    float top[5] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    for (int ch = 0; ch < vchunks; ++ch) epilogue_vocab_chunk(c, p, ch, top);
                    break;
                }
                const int L = ji / 6, k = ji % 6;
                const FusedLayerP& W = p.layer[L];
                if (k == 0) {
                    bf16* cache_t = p.qkv_cache + (static_cast<size_t>(L) * p.T + p.t) * p.R * 3 * FD;
                    for (int ch = 0; ch < 6; ++ch) epilogue_chunk<EPI_CACHE>(c, p, W.b_qkv, ch, 6, bias_reg, cache_t, 3 * FD);
                } else if (k == 1) {
                    epilogue_layernorm<true>(c, p, W.b_o1, W.g1, W.be1, nullptr);
                } else if (k == 2) {
                    for (int ch = 0; ch < 2; ++ch) epilogue_chunk<EPI_CACHE>(c, p, W.b_q, ch, 2, bias_reg, p.q_out, FD);
                } else if (k == 3) {
                    epilogue_layernorm<true>(c, p, W.b_o2, W.g2, W.be2, nullptr);
                } else if (k == 4) {
                    for (int ch = 0; ch < 8; ++ch) epilogue_chunk<EPI_HID>(c, p, W.b_w1, ch, 8, bias_reg, nullptr, 0);
                    fence_proxy_async();  // hidden tile (global, generic proxy) -> TMA reads (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(h_ready);
                } else {
                    epilogue_layernorm<true>(c, p, W.b_w2, W.g3, W.be3, pad_t);  // + zero rows fed <pad> (decoders.py:26)
                }
            }
            pdl_launch_dependents();
        } else {
        const bool tr = (c.ww == 0 && lane == 0);
        fstamp(p, tile, 0, tr);
        pdl_wait();  // tokens / ancestry come from the previous step's selection kernel
        fstamp(p, tile, 1, tr);
        embed_phase(c, p);
        workers_sync(c.nw);  // the residual tile was written warp-per-row, the epilogues read it thread-per-row
        publish_a(c);

        uint8_t* pad_t = p.padflag + static_cast<size_t>(p.t) * p.R;
        for (int L = 0; L < p.n_layers; ++L) {
            const FusedLayerP& W = p.layer[L];
            bf16* cache_l = p.qkv_cache + static_cast<size_t>(L) * p.T * p.R * 3 * FD;
            bf16* cache_t = cache_l + static_cast<size_t>(p.t) * p.R * 3 * FD;
            const int sb = 2 + L * 8;
            fstamp(p, tile, sb, tr);
            for (int ch = 0; ch < 6; ++ch) epilogue_chunk<EPI_CACHE>(c, p, W.b_qkv, ch, 6, bias_reg, cache_t, 3 * FD);
            workers_sync(c.nw);  // q|k|v of every row of the tile are in the cache
            fstamp(p, tile, sb + 1, tr);
            self_attention_phase(c, p, cache_l);
            publish_attention(c);
            fstamp(p, tile, sb + 2, tr);
            epilogue_layernorm<false>(c, p, W.b_o1, W.g1, W.be1, nullptr);
            fstamp(p, tile, sb + 3, tr);
            for (int ch = 0; ch < 2; ++ch) epilogue_chunk<EPI_QG>(c, p, W.b_q, ch, 2, bias_reg, nullptr, 0);
            workers_sync(c.nw);
            fstamp(p, tile, sb + 4, tr);
            cross_attention_phase(c, p, p.cross_kv + static_cast<size_t>(L) * p.cross_layer_stride);
            publish_attention(c);
            fstamp(p, tile, sb + 5, tr);
            epilogue_layernorm<false>(c, p, W.b_o2, W.g2, W.be2, nullptr);
            fstamp(p, tile, sb + 6, tr);
            for (int ch = 0; ch < 8; ++ch) epilogue_chunk<EPI_HID>(c, p, W.b_w1, ch, 8, bias_reg, nullptr, 0);
            fence_proxy_async();  // hidden tile (global, generic proxy) -> bulk-copy reads (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(h_ready);
            fstamp(p, tile, sb + 7, tr);
            epilogue_layernorm<false>(c, p, W.b_w2, W.g3, W.be3, pad_t);  // + zero rows fed <pad> (decoders.py:26)
        }
        fstamp(p, tile, 2 + p.n_layers * 8, tr);
        pdl_launch_dependents();
        const int vchunks = p.vocab_tiles / 2;
        float top[5] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int ch = 0; ch < vchunks; ++ch) epilogue_vocab_chunk(c, p, ch, top);
        fstamp(p, tile, 3 + p.n_layers * 8, tr);
            }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync(); else __syncthreads();   // pair: the leader's MMAs read the peer's shared memory
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
// Stacked bf16 copies of a decoder's projection weights: [layers * 5120][512] (q|k|v, self fc_o, cross fc_q, cross
// fc_o, fc1 per layer) and [layers * 512][2048] (fc2), so that one TMA tensor map serves every GEMM of a chain.
// One set can serve any number of handles (cap_fused_desc::stacked): engines that pipeline independent batches then
// stream the SAME addresses, which the evict_last hint keeps L2-resident.
struct cap_fused_weights {
    void* w512 = nullptr;
    void* w2 = nullptr;
    int n_layers = 0;
};

extern "C" int cap_fused_weights_destroy(cap_fused_weights* w) {
    if (!w) return CAP_OK;
    cudaFree(w->w512);
    cudaFree(w->w2);
    delete w;
    return CAP_OK;
}

extern "C" int cap_fused_weights_create(const cap_fused_layer* layers, int n_layers, cap_fused_weights** out) {
    CAP_REQUIRE(layers && out, "cap_fused_weights_create: null pointer");
    CAP_REQUIRE(n_layers >= 1 && n_layers <= MAX_FUSED_LAYERS, "cap_fused_weights_create: 1..%d layers", MAX_FUSED_LAYERS);
    cap_fused_weights* f = new cap_fused_weights();
    f->n_layers = n_layers;
    auto fail = [&](int rc) { cap_fused_weights_destroy(f); return rc; };
    const size_t w512_elems = static_cast<size_t>(n_layers) * W512_ROWS_PER_LAYER * FD;
    const size_t w2_elems = static_cast<size_t>(n_layers) * FD * FDFF;
    if (cudaMalloc(&f->w512, w512_elems * 2) != cudaSuccess || cudaMalloc(&f->w2, w2_elems * 2) != cudaSuccess)
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: cudaMalloc of the stacked weights failed"));
    for (int l = 0; l < n_layers; ++l) {
        const cap_fused_layer& w = layers[l];
        const void* srcs[5] = {w.w_qkv, w.w_o1, w.w_q, w.w_o2, w.w_fc1};
        const int rows[5] = {3 * FD, FD, FD, FD, FDFF};
        bf16* dst = static_cast<bf16*>(f->w512) + static_cast<size_t>(l) * W512_ROWS_PER_LAYER * FD;
        for (int i = 0; i < 5; ++i) {
            if (!srcs[i]) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_weights_create: null weight"));
            if (cudaMemcpy(dst, srcs[i], static_cast<size_t>(rows[i]) * FD * 2, cudaMemcpyDeviceToDevice) != cudaSuccess)
                return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: weight copy failed"));
            dst += static_cast<size_t>(rows[i]) * FD;
        }
        if (!w.w_fc2 || cudaMemcpy(static_cast<bf16*>(f->w2) + static_cast<size_t>(l) * FD * FDFF, w.w_fc2,
                                   static_cast<size_t>(FD) * FDFF * 2, cudaMemcpyDeviceToDevice) != cudaSuccess)
            return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: fc2 copy failed"));
    }
    *out = f;
    return CAP_OK;
}

struct cap_fused_decoder {
    FusedParams base;
    cap_fused_weights* stacked = nullptr;   // the stacked weights the tensor maps point into
    bool owns_stacked = false;              // false: cap_fused_desc::stacked, owned by the caller
    int tiles = 0;
    bool has_att = false;
    bool use_pairs = true;      // CTA pairs (OPENVIIC_CHAIN_PAIR=0 at creation: single CTAs)
    bool full_logits = false;   // debug / parity: every logit is stored (cap_fused_set_full_logits, OPENVIIC_FULL_LOGITS)
};

extern "C" int cap_fused_create(const cap_fused_desc* d, cap_fused_decoder** out) {
    CAP_REQUIRE(d && out, "cap_fused_create: null pointer");
    CAP_REQUIRE(d->d_model == FD && d->heads == FHEADS && d->d_ff == FDFF, "cap_fused_create: needs d_model 512, 8 heads, d_ff 2048");
    CAP_REQUIRE(d->n_layers >= 1 && d->n_layers <= MAX_FUSED_LAYERS, "cap_fused_create: 1..%d layers", MAX_FUSED_LAYERS);
    CAP_REQUIRE(d->beam >= 1 && d->beam <= MAXB, "cap_fused_create: beam must be 1..%d", MAXB);
    CAP_REQUIRE(d->max_rows > 0 && d->vocab > 8 && d->max_len > 0 && d->max_len <= 40, "cap_fused_create: bad sizes (max_len <= 40)");
    CAP_REQUIRE(d->ld_logits % 32 == 0 && d->ld_logits >= d->vocab, "cap_fused_create: ld_logits must be a multiple of 32");
    CAP_REQUIRE(d->stacked == nullptr || d->stacked->n_layers == d->n_layers, "cap_fused_create: stacked weights of another model");
    CAP_PROPAGATE(install_fault_buffer());
    cap_fused_decoder* f = new cap_fused_decoder();
    FusedParams& p = f->base;
    memset(&p, 0, sizeof(p));
    const int L = d->n_layers;
    p.n_layers = L;
    f->tiles = ((d->max_rows + TILE_ROWS - 1) / TILE_ROWS + 1) / 2 * 2;   // even: CTA pairs; scratch tiles exist for a dummy partner
    auto fail = [&](int rc) { cap_fused_destroy(f); return rc; };
    if (d->stacked) {
        f->stacked = const_cast<cap_fused_weights*>(d->stacked);
    } else {
        const int rc0 = cap_fused_weights_create(d->layers, L, &f->stacked);
        if (rc0 != CAP_OK) return fail(rc0);
        f->owns_stacked = true;
    }
    for (int l = 0; l < L; ++l) {
        const cap_fused_layer& w = d->layers[l];
        FusedLayerP& lp = p.layer[l];
        lp.b_qkv = w.b_qkv; lp.b_o1 = w.b_o1; lp.g1 = w.ln1_g; lp.be1 = w.ln1_b;
        lp.b_q = w.b_q; lp.b_o2 = w.b_o2; lp.g2 = w.ln2_g; lp.be2 = w.ln2_b;
        lp.b_w1 = w.b_fc1; lp.b_w2 = w.b_fc2; lp.g3 = w.ln3_g; lp.be3 = w.ln3_b;
        const float* need[12] = {lp.b_qkv, lp.b_o1, lp.g1, lp.be1, lp.b_q, lp.b_o2, lp.g2, lp.be2, lp.b_w1, lp.b_w2, lp.g3, lp.be3};
        for (const float* q : need)
            if (!q) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_create: null bias / LayerNorm parameter"));
    }
    void* const w512 = f->stacked->w512;
    void* const w2 = f->stacked->w2;
    int rc = cap_gemm::make_tmap(&p.map_w512, w512, L * W512_ROWS_PER_LAYER, FD, FD, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_w2, w2, L * FD, FDFF, FDFF, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_vocab, d->w_vocab, d->vocab, FD, FD, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_w512_h, w512, L * W512_ROWS_PER_LAYER, FD, FD, 64);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_w2_h, w2, L * FD, FDFF, FDFF, 64);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_vocab_h, d->w_vocab, d->vocab, FD, FD, 64);
    if (rc != CAP_OK) return fail(rc);
    p.tokens = d->tokens; p.word_emb = static_cast<const bf16*>(d->word_emb); p.word_pos = d->word_pos; p.pad_idx = d->pad_idx;
    p.qkv_cache = static_cast<bf16*>(d->qkv_cache); p.ancestry = d->ancestry; p.padflag = d->padflag;
    p.cross_kv = static_cast<const bf16*>(d->cross_kv); p.cross_layer_stride = d->cross_layer_stride; p.enc_mask = d->enc_mask;
    p.logits = d->logits; p.ld_logits = d->ld_logits; p.part_ms = d->part_ms;
    p.vocab = d->vocab;
    p.vocab_tiles = ((d->vocab + 255) / 256) * 2;
    p.stat_chunks = ((d->vocab + 255) / 256) * 8;
    p.T = d->max_len; p.beam = d->beam;
    p.scale = 1.0f / sqrtf(static_cast<float>(FD / FHEADS));
    const size_t tiles = f->tiles;
    void *res = nullptr, *qg = nullptr, *hb = nullptr;
    if (cudaMalloc(&res, tiles * TILE_ROWS * FD * 4) != cudaSuccess || cudaMalloc(&qg, tiles * TILE_ROWS * FD * 2) != cudaSuccess ||
        cudaMalloc(&hb, tiles * TILE_ROWS * FDFF * 2) != cudaSuccess) {
        cudaFree(res); cudaFree(qg); cudaFree(hb);
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_create: cudaMalloc of the scratch tiles failed"));
    }
    cudaMemset(res, 0, tiles * TILE_ROWS * FD * 4);
    cudaMemset(qg, 0, tiles * TILE_ROWS * FD * 2);
    cudaMemset(hb, 0, tiles * TILE_ROWS * FDFF * 2);
    p.res = static_cast<float*>(res); p.qg = static_cast<bf16*>(qg); p.hbuf = static_cast<bf16*>(hb);
    rc = cap_gemm::make_tmap(&p.map_h, hb, static_cast<int>(tiles) * TILE_ROWS, FDFF, FDFF, 128);
    if (rc != CAP_OK) return fail(rc);
    if (d->att_in) {
        rc = cap_gemm::make_tmap(&p.map_att, d->att_in, d->max_rows, FD, FD, 128);
        if (rc != CAP_OK) return fail(rc);
    }
    p.q_out = static_cast<bf16*>(d->q_out);
    f->has_att = d->att_in != nullptr;
    f->full_logits = getenv("OPENVIIC_FULL_LOGITS") && atoi(getenv("OPENVIIC_FULL_LOGITS")) != 0;
    f->use_pairs = !(getenv("OPENVIIC_CHAIN_PAIR") && atoi(getenv("OPENVIIC_CHAIN_PAIR")) == 0);
    if (cudaFuncSetAttribute(decode_step_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(decode_step_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(decode_step_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess)
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_create: cannot reserve %u bytes of shared memory", FUSED_SMEM));
    *out = f;
    return CAP_OK;
}

extern "C" int cap_fused_set_full_logits(cap_fused_decoder* f, int on) {
    CAP_REQUIRE(f != nullptr, "cap_fused_set_full_logits: null handle");
    f->full_logits = on != 0;
    return CAP_OK;
}

extern "C" int cap_fused_get_full_logits(cap_fused_decoder* f) { return f && f->full_logits ? 1 : 0; }

extern "C" int cap_fused_destroy(cap_fused_decoder* f) {
    if (!f) return CAP_OK;
    if (f->owns_stacked) cap_fused_weights_destroy(f->stacked);
    cudaFree(f->base.res);
    cudaFree(f->base.qg);
    cudaFree(f->base.hbuf);
    delete f;
    return CAP_OK;
}

static unsigned long long* g_fused_trace = nullptr;
extern "C" int cap_debug_fused_trace(unsigned long long* device_buffer) {
    g_fused_trace = device_buffer;
    return CAP_OK;
}

extern "C" int cap_fused_decode_step(cap_fused_decoder* f, int t, int B, int n_keys, cap_stream_t stream) {
    CAP_REQUIRE(f != nullptr, "cap_fused_decode_step: null handle");
    FusedParams p = f->base;
    CAP_REQUIRE(t >= 0 && t < p.T, "cap_fused_decode_step: step %d outside [0,%d)", t, p.T);
    const int R = B * p.beam;
    const int tiles = (R + TILE_ROWS - 1) / TILE_ROWS;
    CAP_REQUIRE(B > 0 && tiles <= f->tiles && n_keys > 0 && n_keys <= 128, "cap_fused_decode_step: batch %d / %d keys unsupported", B, n_keys);
    p.t = t; p.R = R; p.B = B; p.n_keys = n_keys;
    p.trace = g_fused_trace;
    static const int dbg_skip = getenv("OPENVIIC_FUSED_DBG_SKIP") ? atoi(getenv("OPENVIIC_FUSED_DBG_SKIP")) : 0;
    p.dbg_skip = dbg_skip;
    // plain launch (no programmatic dependent launch), for the reason given in cap_fused_chain below
    decode_step_fused_kernel<false><<<dim3(tiles), dim3(FUSED_THREADS), FUSED_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("decode_step_fused_kernel");
}

extern "C" int cap_fused_chain(cap_fused_decoder* f, int chain, int layer, int t, int B, cap_stream_t stream) {
    CAP_REQUIRE(f != nullptr, "cap_fused_chain: null handle");
    FusedParams p = f->base;
    CAP_REQUIRE(t >= 0 && t < p.T, "cap_fused_chain: step %d outside [0,%d)", t, p.T);
    CAP_REQUIRE(chain >= CAP_CHAIN_EMBED_QKV && chain <= CAP_CHAIN_FFN && layer >= 0 && layer < p.n_layers,
                "cap_fused_chain: bad chain %d / layer %d", chain, layer);
    CAP_REQUIRE(p.q_out != nullptr && (chain == CAP_CHAIN_EMBED_QKV || f->has_att), "cap_fused_chain: handle has no att_in / q_out buffers");
    const int R = B * p.beam;
    const int tiles = (R + TILE_ROWS - 1) / TILE_ROWS;
    CAP_REQUIRE(B > 0 && tiles <= f->tiles, "cap_fused_chain: batch %d exceeds the reservation", B);
    p.t = t; p.R = R; p.B = B; p.n_keys = 0;
    p.sparse_logits = (f->full_logits || p.beam > 5) ? 0 : 1;
    p.trace = g_fused_trace;   // debug (cap_debug_fused_trace): the epilogue stamps of fc1's chunk 3, else nullptr
    p.dbg_skip = 0;
    p.start_embed = 0;
    if (chain == CAP_CHAIN_EMBED_QKV) {          // x = Emb + pos; q|k|v of layer 0 -> cache
        CAP_REQUIRE(layer == 0, "cap_fused_chain: the embedding chain belongs to layer 0");
        p.start_embed = 1; p.job_begin = 0; p.job_end = 0;
    } else if (chain == CAP_CHAIN_SELF_OUT) {    // self fc_o + LN; cross fc_q -> q_out
        p.job_begin = layer * 6 + 1; p.job_end = layer * 6 + 2;
    } else {                                     // cross fc_o + LN; FFN + LN; next layer's q|k|v or the vocabulary
        p.job_begin = layer * 6 + 3; p.job_end = layer * 6 + 6;
    }
    // No programmatic dependent launch for a chain: its CTAs each take a whole SM, and an early-started chain
    // would hold them idle in griddepcontrol.wait until the attention kernel before it has drained.
    // CTA pairs (OPENVIIC_CHAIN_PAIR=0: single CTAs): clusters of two adjacent tiles, the second CTA of the last
    // pair is a dummy when the tile count is odd (every access of it is guarded by the row count).
    if (f->use_pairs) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((tiles + 1) / 2 * 2);
        cfg.blockDim = dim3(CHAIN_THREADS);
        cfg.dynamicSmemBytes = FUSED_SMEM;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, decode_step_fused_kernel<true, true>, p);
    } else {
        decode_step_fused_kernel<true><<<dim3(tiles), dim3(CHAIN_THREADS), FUSED_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
    }
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("decode_step_fused_kernel<chain>");
}
