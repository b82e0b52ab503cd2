// GEMM chains of a caption decode step on tcgen05: a CTA (or a pair of CTAs) owns a tile of 128 beam rows and carries
// it through a LIST OF JOBS -- projection GEMMs with their epilogues (bias / ReLU / residual + LayerNorm / meshed gates
// / log-softmax chunk statistics) -- without the activations leaving the SM between them.  Rows never interact inside
// a step (beams interact only in the selection, beam.cu), so there is no grid-wide dependency: nothing but the weights
// is shared between CTAs.  The attention kernels (attention.cu) run between the chains.
//
// Reference operators covered by the job lists (cap_fused_chain builds them):
//   token embedding                        decoders.py:105-112
//   q|k|v, fc_o + residual + LayerNorm     attentions.py:44-58, 296-309
//   cross fc_q                             attentions.py:51 (queries of decoders.py:23,56)
//   meshed gates + mix                     decoders.py:55-67   (MeshedDecoderLayer)
//   fc1 -> ReLU -> fc2 + residual + LN     positionwise_feed_forward.py:23-28, zero rows fed <pad> decoders.py:26
//   vocabulary projection (+ statistics)   decoders.py:90,121-123
//
// Structure of a CTA (384 threads; setmaxnreg moves registers from the control warpgroup to the workers):
//   warp 0   : producer -- streams weight tiles (TMA, SWIZZLE_128B) through a ring for ALL the chain's GEMMs back to
//              back: weights do not depend on activations, so the ring is refilled while the workers are in an epilogue;
//   warp 1   : TMEM allocator + tcgen05.mma issuer (UMMA 128x128x16 / 256x128x16 for CTA pairs, fp32 accumulators in
//              two 256-column TMEM buffers: the epilogue of one 256-column chunk overlaps the MMAs of the next);
//   warp 2   : A-tile loader -- attention outputs arrive by TMA (the chain's first tile, and mid-chain tiles of the
//              meshed decoder's encoder levels);
//   warps 4-11: workers -- the epilogues and the embedding.
// The activation tile is the RESIDENT A operand: 128 rows x 512 columns of bf16 in shared memory as eight K-major
// SWIZZLE_128B k-blocks (the layout TMA would produce), written directly by the epilogues: a thread per row puts its
// 16-byte chunk c of k-block kb at kb*16K + row*128 + ((c ^ (row & 7)) << 4), which is bank-conflict-free for a
// thread-per-row writer and is what the tcgen05.mma descriptor expects.  Only the 2048-wide FFN hidden tile does not
// fit: it goes through a row-major global scratch buffer and comes back through TMA as the streamed A operand of the
// second FFN GEMM.  The fp32 residual stream (and the meshed decoder's fp32 side streams) live in global scratch
// buffers in 16-byte-granule layout [granule][row] (coalesced for a thread-per-row reader; L2-resident).
// CTA pairs: two CTAs of a cluster, each with its own 128-row tile, share every weight tile -- each stages HALF of a
// 256-row tile pair (128 rows), the leader issues tcgen05.mma.cta_group::2 (M = 256, N = 256) over both shared
// memories and both CTAs drain their own accumulator rows: per CTA the same 64 KB ring feeds twice the tensor work.
//
// (Removed in round 2, after measurement: the one-kernel-per-step mode that also ran both attention phases on the
// CTA's CUDA cores -- parity-green but 2.4x slower per tile than chains + stand-alone attention, because a chain CTA
// owns its SM and its low-IPC attention phases idled the tensor pipe; and two epilogue variants, see epilogue_store.)
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
constexpr int NW = 8;                // worker warps (16 were measured no faster: the chains are paced by the weight ring)
constexpr int FIRST_WORKER_WARP = 4;  // warps 0-3 form the control warpgroup (producer, MMA issuer, A loader, one idle)
constexpr int CHAIN_THREADS = (FIRST_WORKER_WARP + NW) * 32;
constexpr int CHAIN_MAXNREG = 128;
constexpr int NB = 4;                // weight ring stages (single CTAs)
constexpr uint32_t B_STAGE_BYTES = 128 * 64 * 2;   // one 128-row x 64-column weight tile
constexpr uint32_t A_KB_BYTES = 128 * 64 * 2;      // one k-block of the A operand
constexpr int A_SLOTS = 8;           // k-blocks of the resident A tile = slots of the streamed-A ring
constexpr int STAGE_PITCH = 80;      // per-warp store staging: 32 rows x 64 B (+16 B skew)
constexpr int MAXB = 5;              // beams per image the sparse-logits bookkeeping is written for
constexpr int MAX_FUSED_LAYERS = 6;
constexpr int MAX_LEVELS = 3;        // encoder levels of the meshed decoder
constexpr int MAX_JOBS = 12;         // jobs of one chain launch

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + A_SLOTS * A_KB_BYTES;
constexpr uint32_t OFF_STAGE = OFF_B + NB * B_STAGE_BYTES;
constexpr uint32_t STAGE_BYTES_ALL = NW * 32 * STAGE_PITCH;
constexpr uint32_t OFF_BIAS = OFF_STAGE + STAGE_BYTES_ALL;
constexpr uint32_t OFF_CBIAS = OFF_BIAS + FD * 4;     // [2][256] chunk biases of the plain projections
constexpr uint32_t OFF_GAMMA = OFF_CBIAS + FD * 4;
constexpr uint32_t OFF_BETA = OFF_GAMMA + FD * 4;
// LayerNorm row statistics [2 (sum, sumsq)][2 halves][128 rows] alias the staging tiles: the LN epilogue stages
// nothing, and a workers_sync separates it from the chunk epilogues on both sides
constexpr uint32_t OFF_STAT = OFF_STAGE;
static_assert(2 * 2 * TILE_ROWS * 4 <= STAGE_BYTES_ALL, "row statistics must fit the staging area");
constexpr uint32_t OFF_BARS = OFF_BETA + FD * 4;
// CTA pairs issue tcgen05.mma with N = 256 (two adjacent 128-column tiles per instruction): each CTA stages its 128 of
// the 256 weight rows of a k-block, 16 KB per stage like a single CTA.  (Round 2 first ran pairs with N = 128 and 64-row
// half tiles, 8 stages of 8 KB: the issuer warp then needed ~370 cycles of descriptor arithmetic, R2UR moves and barrier
// traffic per stage for 256 cycles of tensor work -- ncu showed the tensor pipe active 45 % of an FFN chain while the
// issuer accounted 66 % of its time to issuing.  Twice the work per issued instruction puts the pipe back in front.)
constexpr int NB_PAIR = 4;
constexpr uint32_t B_STAGE_BYTES_PAIR = 128 * 64 * 2;
constexpr int TILE_STEP_PAIR = 2;     // 128-column tiles per MMA instruction
constexpr int NUM_BARS = 2 * NB_PAIR + 2 * A_SLOTS + 2 + 2 + 4;
constexpr uint32_t OFF_TMEM = OFF_BARS + NUM_BARS * 8;
constexpr uint32_t FUSED_SMEM = OFF_TMEM + 16;

// ---- jobs ---------------------------------------------------------------------------------------------------
enum JobEpi : int {
    JE_STORE = 0,   // + bias (+ ReLU) -> bf16 rows at dst, 256-column chunks
    JE_LN = 1,      // N = 512: + bias + residual (res_in) -> LayerNorm -> resident A tile (bf16) + res_out (fp32) [zero rows]
    JE_VOCAB = 2,   // bias-free: fp32 logits (sparse) + per-32-column (max, sum exp) statistics
    JE_F32 = 3,     // raw fp32 accumulators -> side stream `aux` in granule layout (the meshed gates' s-part)
    JE_ALPHA = 4    // N = 512: sigmoid(acc + bias + aux[level]) * c (res_in) accumulated into the mix stream (res_out);
                    //          last level: scaled by 1/sqrt(levels) -> resident A tile + residual stream
};
enum JobASrc : int {
    JA_RESIDENT = 0,         // the resident tile, written by the previous epilogue or the embedding
    JA_TMA_TILE = 1,         // a tile of attention output arrives by TMA (att_in)
    JA_STREAM_HIDDEN = 2,    // K = 2048: the FFN hidden tile streamed back from its scratch buffer
    JA_STREAM_FEATURES = 3   // K = d_feature: the raw visual features streamed from global memory (encoder's first GEMM)
};
enum JobFlags : int {
    JF_RELU = 1, JF_WHOLE_TILES = 2, JF_HIDDEN_DONE = 4, JF_LAST_LEVEL = 8, JF_FIRST_LEVEL = 16,
    JF_WAIT_A = 32,  // the resident A tile is (re)written just before this job: the issuer waits for the workers' publish
    JF_NO_RESIDUAL = 64   // JE_LN: LayerNorm of the projection alone (vision embedding, encoders.py:36)
};

struct JobDesc {
    int wmap;        // weight tensor map: 0 K = 512 stack, 1 fc2 stack (K = 2048), 2 vocabulary, 3 vision projection (K = d_feature)
    int row0;        // first weight row of the job inside that stack
    int ntiles;      // 128-row weight tiles (N / 128)
    int kblocks;     // K / 64
    int chunk;       // n-tiles per accumulator pass: 2 (one 256-column TMEM buffer) or 4 (both)
    int a_src;       // JobASrc
    int att_row0;    // JA_TMA_TILE: first row of this job's tile source in the att_in tensor map (level * max_rows)
    int epi;         // JobEpi
    int flags;       // JobFlags
    int aux_col0;    // JE_ALPHA: first column of this level's gate inside the aux stream
    int ld_dst;      // JE_STORE: destination row pitch (elements)
    int pad_;
    const float* bias;
    const float* gamma;
    const float* beta;
    bf16* dst;               // JE_STORE
    const float* res_in;     // JE_LN residual / JE_ALPHA c stream   (granule layout, per tile)
    float* res_out;          // JE_LN output stream / JE_ALPHA mix or final stream
    const uint8_t* zero_rows;  // JE_LN: rows to zero (fed <pad> / padded visual tokens), or nullptr
    const float* pos;          // JE_LN: position table [pos_rows][512] added after the LayerNorm (row % pos_rows), or nullptr
    bf16* ln_out;              // JE_LN: row-major bf16 copy of the result [R][512] (encoder level outputs), or nullptr
};

struct FusedParams {
    CUtensorMap map_w512;   // [layers * rows_per_layer][512]: the model's K = 512 projections stacked
    CUtensorMap map_w2;     // [layers * 512][2048]
    CUtensorMap map_vocab;  // [V][512]
    CUtensorMap map_h;      // [tiles * 128][2048] FFN hidden scratch
    CUtensorMap map_att;    // [levels * max_rows][512] attention outputs of the stand-alone attention kernels
    CUtensorMap map_wvis;   // encoder: vision projection [512][d_feature]
    CUtensorMap map_feat;    // encoder: raw features [rows][d_feature], the streamed A operand of the vision projection
    int pos_rows;            // encoder: visual tokens per image (rows of the position table)
    JobDesc jobs[MAX_JOBS];
    int n_jobs;
    int start_embed;         // the chain starts with x = Emb[token] + pos (its first job reads the resident tile)
    const int32_t* tokens;
    const bf16* word_emb;
    const float* word_pos;
    int pad_idx;
    uint8_t* padflag;        // [T][R]
    float* res;              // [tiles][128 granules of 4 floats][128 rows][4]   fp32 residual stream
    float* aux;              // [tiles][aux_cols / 4][128 rows][4]               fp32 side stream (meshed gates, s-part)
    int aux_cols;
    bf16* hbuf;              // [tiles * 128][2048] row-major                    FFN hidden
    float* logits;
    int ld_logits;
    float* part_ms;          // [R][stat_chunks][2]
    int vocab, vocab_tiles, stat_chunks;
    int t, T, R, B, beam;
    float level_scale;       // 1 / sqrt(levels) of the meshed mix
    int sparse_logits;       // 1: store only the 32-column groups that can hold one of the row's top-5 candidates
    unsigned long long* trace;  // debug: %globaltimer stamps per CTA (cap_debug_fused_trace), else nullptr
};

__device__ __forceinline__ void fstamp(const FusedParams& p, int tile, int slot, bool who) {
    if (p.trace != nullptr && who) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[static_cast<size_t>(tile) * 64 + slot] = t;
    }
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

__device__ __forceinline__ uint4 pack8_u4(const float* f) {
    bf16x8 p = pack8(f);
    return *reinterpret_cast<uint4*>(&p);
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }

struct WorkerCtx {
    uint8_t* A_buf;
    uint8_t* stage;   // this warp's staging tile
    float *s_bias, *s_cbias, *s_gamma, *s_beta, *s_stat;
    uint64_t *acc_full, *acc_empty, *a_ready, *h_ready;
    bool pair_peer;    // CTA pair, and this CTA is not the leader: consumer-side barriers live in the leader CTA
    uint32_t tmem_base;
    uint32_t use0, use1;  // completed uses of TMEM buffer 0 / 1 (scalars: no dynamically indexed state)
    int toggle;
    int ww, quad, half, lane, wtid;   // half = ww / 4: which 128 (of 256) / 256 (of 512) columns of a quadrant's rows
    int tile, r0, rows_valid_warp;    // rows of this warp's TMEM quadrant that exist (0..32)
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

// Epilogue of one 256-column chunk of a plain projection: + bias (ReLU for the hidden layer), bf16, to global rows.
// `bias_reg` carries this thread's bias value of the chunk across calls: the value of chunk c + 1 is requested
// while chunk c is drained, so the global-load latency is off the per-chunk critical path.
// The four tcgen05.ld of a warp run one group ahead of the math.  (First measured while the MMA issuer was the
// bottleneck, where it changed nothing -- 84.5 k vs 85.2 k captions/s; with N = 256 instructions the chains wait for
// free accumulators instead, and the epilogue is on the critical path.  Storing the bf16 rows straight from registers
// instead of through the staging tile was slower, 81.9 k: the store sectors cost more than the shared-memory round
// trip.)
__device__ __forceinline__ void epilogue_store(WorkerCtx& c, const FusedParams& p, const JobDesc& job, int chunk_idx,
                                               int n_chunks, float& bias_reg) {
    const int flags = job.flags;   // job fields are read once: the asm wrappers' memory clobbers would reload them
    const float* bias = job.bias;
    bf16* const dst = job.dst;
    const int ld_dst = job.ld_dst;
    const bool tr = ((flags & JF_RELU) && chunk_idx == 3 && c.ww == 0 && c.lane == 0);
    fstamp(p, c.tile, 40, tr);
    float* sb = c.s_cbias + (chunk_idx & 1) * 256;   // double-buffered by chunk parity (see the barrier note below)
    if (c.wtid < 256) {
        if (chunk_idx == 0) bias_reg = __ldg(bias + c.wtid);
        sb[c.wtid] = bias_reg;
        if (chunk_idx + 1 < n_chunks) bias_reg = __ldg(bias + (chunk_idx + 1) * 256 + c.wtid);
    }
    workers_sync();  // a warp reaches the NEXT chunk's barrier only after it has finished reading this one
    fstamp(p, c.tile, 41, tr);
    const int b = acquire_acc(c, 2);
    fstamp(p, c.tile, 42, tr);
    const float relu_floor = (flags & JF_RELU) ? 0.f : -INFINITY;   // one FMNMX either way, no branch in the loop
    const int rows_valid = (flags & JF_WHOLE_TILES) ? 32 : c.rows_valid_warp;   // scratch buffers hold whole tiles
    uint8_t* const tile_base = reinterpret_cast<uint8_t*>(dst + static_cast<size_t>(c.r0 + c.quad * 32) * ld_dst + chunk_idx * 256);
    // the tcgen05.ld of group i + 1 is in flight while group i is converted and stored
    const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16) + b * 256 + c.half * 128;
    uint32_t vv[2][32];
    tmem_ld_32x32b_x32(taddr, vv[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int colc = (c.half * 4 + i) * 32;
        uint32_t* v = vv[i & 1];
        tmem_ld_wait_on(v);
        if (i + 1 < 4) tmem_ld_32x32b_x32(taddr + (i + 1) * 32, vv[(i + 1) & 1]);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            f[j] = fmaxf(__uint_as_float(v[j]) + sb[colc + j], relu_floor);
        }
        uint4 o[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) o[g] = pack8_u4(f + 8 * g);
        staged_store64(c.stage, c.lane, o, tile_base + colc * 2, static_cast<size_t>(ld_dst) * 2, rows_valid);
        fstamp(p, c.tile, 43 + i, tr);
    }
    release_acc(c, 2, b);
    fstamp(p, c.tile, 47, tr);
}

// Epilogue of one 256-column chunk whose raw fp32 accumulators go to the side stream (granule layout): the s-part of
// the meshed decoder's gates, s . W_alpha_i[:, :512]^T for the three levels (decoders.py:60-62), computed while the
// self-attention output s is the resident tile and consumed by JE_ALPHA of the same layer.
__device__ __forceinline__ void epilogue_f32(WorkerCtx& c, const FusedParams& p, int chunk_idx) {
    const int b = acquire_acc(c, 2);
    const int row = c.quad * 32 + c.lane;
    float4* aux = reinterpret_cast<float4*>(p.aux) + static_cast<size_t>(c.tile) * (p.aux_cols / 4) * TILE_ROWS + row;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int colc = (c.half * 4 + i) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16) + b * 256 + colc, v);
        tmem_ld_wait();
        const int gcol = chunk_idx * 256 + colc;
#pragma unroll
        for (int g = 0; g < 8; ++g)
            aux[static_cast<size_t>(gcol / 4 + g) * TILE_ROWS] =
                make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                            __uint_as_float(v[4 * g + 3]));
    }
    release_acc(c, 2, b);
}

// Epilogue of an N = 512 projection followed by residual + LayerNorm (attentions.py:308-309,
// positionwise_feed_forward.py:26): thread = row, the two warps of a TMEM quadrant take 256 columns each and
// exchange (sum, sum of squares); the normalised row goes to the resident A tile (bf16) and to the fp32 stream
// res_out (which may be res_in: a thread only ever touches its own elements).  Pass A writes y = projection + bias +
// residual back into the accumulator's TMEM columns, so pass B needs neither the residual (global) nor the bias
// again; the residual granules of the next 32 columns are requested while the current TMEM load is in flight.
__device__ __forceinline__ void epilogue_layernorm(WorkerCtx& c, const FusedParams& p, const JobDesc& job) {
    // every warp has left the previous epilogue: a LayerNorm / gate epilogue may still be reading s_bias / s_gamma /
    // s_beta in another quadrant (its quadrant barrier joins two warps only), and s_stat aliases the staging tiles
    workers_sync();
    for (int i = c.wtid; i < FD; i += NW * 32) {
        c.s_bias[i] = __ldg(job.bias + i);
        c.s_gamma[i] = __ldg(job.gamma + i);
        c.s_beta[i] = __ldg(job.beta + i);
    }
    workers_sync();
    acquire_acc(c, 4);
    const int row = c.quad * 32 + c.lane;
    const int grow = c.r0 + row;
    const bool live = grow < p.R;
    const size_t tile_off = static_cast<size_t>(c.tile) * (FD / 4) * TILE_ROWS + row;
    const bool has_res = (job.flags & JF_NO_RESIDUAL) == 0;
    const float4* res_in = reinterpret_cast<const float4*>(job.res_in) + tile_off;
    float4* res_out = reinterpret_cast<float4*>(job.res_out) + tile_off;
    const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16);
    float s1 = 0.f, s2 = 0.f;
    {
        // the residual granules of group i + 1 are requested while group i is summed and parked, and pass B requests the
        // tcgen05.ld of group i + 1 as soon as group i has been normalised (the epilogue is a serial point of the chain --
        // no GEMM of this tile can start before it ends -- so its latency is the chain's latency; double-buffering the
        // accumulator registers as the store epilogues do spills at 128 registers)
        const int cbase = c.half * 256;
        float4 r[8];
#pragma unroll
        for (int g = 0; g < 8; ++g)
            r[g] = has_res ? res_in[static_cast<size_t>(cbase / 4 + g) * TILE_ROWS] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const int c0 = cbase + i * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + c0, v);
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
            if (i + 1 < 8 && has_res) {
#pragma unroll
                for (int g = 0; g < 8; ++g) r[g] = res_in[static_cast<size_t>((c0 + 32) / 4 + g) * TILE_ROWS];
            }
            tmem_st_32x32b_x32(taddr + c0, v);
        }
    }
    tmem_st_wait();
    c.s_stat[c.half * TILE_ROWS + row] = s1;
    c.s_stat[(2 + c.half) * TILE_ROWS + row] = s2;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + c.quad) : "memory");  // the two warps of the quadrant
    const float S1 = c.s_stat[row] + c.s_stat[TILE_ROWS + row];
    const float S2 = c.s_stat[2 * TILE_ROWS + row] + c.s_stat[3 * TILE_ROWS + row];
    const float mean = S1 * (1.f / FD);
    const float var = fmaxf(S2 * (1.f / FD) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const bool zero = !live || (job.zero_rows != nullptr && job.zero_rows[grow] != 0);
    // encoder extras: + pos[row % tokens] after the LayerNorm (encoders.py:36), and a row-major bf16 copy of the result
    const float4* pos = job.pos != nullptr ? reinterpret_cast<const float4*>(job.pos + static_cast<size_t>(live ? grow % p.pos_rows : 0) * FD) : nullptr;
    uint4* ln_out = (job.ln_out != nullptr && live) ? reinterpret_cast<uint4*>(job.ln_out + static_cast<size_t>(grow) * FD) : nullptr;
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + c.half * 256, v);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int c0 = (c.half * 8 + i) * 32;
        tmem_ld_wait_on(v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
            f[j] = zero ? 0.f : (__uint_as_float(v[j]) - mean) * rstd * c.s_gamma[c0 + j] + c.s_beta[c0 + j];
        if (i + 1 < 8) tmem_ld_32x32b_x32(taddr + c0 + 32, v);   // v is dead: the next group loads while this one is stored
        if (pos != nullptr && !zero) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 pp = __ldg(pos + c0 / 4 + g);
                f[4 * g] += pp.x; f[4 * g + 1] += pp.y; f[4 * g + 2] += pp.z; f[4 * g + 3] += pp.w;
            }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
            res_out[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 packed = pack8_u4(f + 8 * g);
            *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row, c0 / 8 + g)) = packed;
            if (ln_out != nullptr) ln_out[c0 / 8 + g] = packed;
        }
    }
    release_acc(c, 4, 0);
    publish_a(c);
}

// Epilogue of the meshed decoder's gate of one encoder level (decoders.py:60-67): the accumulator holds
// c_i . W_alpha_i[:, 512:]^T; alpha_i = sigmoid(acc + s-part (aux stream) + bias); mix += alpha_i * c_i, all in fp32,
// thread = row.  The last level scales the mix by 1/sqrt(levels) and hands it on as the resident A tile (bf16) and the
// residual stream (fp32) of the feed-forward block.
__device__ __forceinline__ void epilogue_alpha(WorkerCtx& c, const FusedParams& p, const JobDesc& job, float* mix_stream) {
    workers_sync();   // the LayerNorm epilogue before this one reads s_bias until its last warp is through pass A
    for (int i = c.wtid; i < FD; i += NW * 32) c.s_bias[i] = __ldg(job.bias + i);
    workers_sync();
    acquire_acc(c, 4);
    const int row = c.quad * 32 + c.lane;
    const bool live = c.r0 + row < p.R;
    const bool first = (job.flags & JF_FIRST_LEVEL) != 0, last = (job.flags & JF_LAST_LEVEL) != 0;
    const size_t tile_off = static_cast<size_t>(c.tile) * (FD / 4) * TILE_ROWS + row;
    const float4* cs = reinterpret_cast<const float4*>(job.res_in) + tile_off;
    float4* mix = reinterpret_cast<float4*>(mix_stream) + tile_off;
    float4* out = reinterpret_cast<float4*>(job.res_out) + tile_off;
    const float4* gate = reinterpret_cast<const float4*>(p.aux) + static_cast<size_t>(c.tile) * (p.aux_cols / 4) * TILE_ROWS + row +
                         static_cast<size_t>(job.aux_col0 / 4) * TILE_ROWS;
    const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int c0 = (c.half * 8 + i) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        float4 r[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) r[g] = gate[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float x0 = __uint_as_float(v[4 * g]) + c.s_bias[c0 + 4 * g] + r[g].x;
            const float x1 = __uint_as_float(v[4 * g + 1]) + c.s_bias[c0 + 4 * g + 1] + r[g].y;
            const float x2 = __uint_as_float(v[4 * g + 2]) + c.s_bias[c0 + 4 * g + 2] + r[g].z;
            const float x3 = __uint_as_float(v[4 * g + 3]) + c.s_bias[c0 + 4 * g + 3] + r[g].w;
            f[4 * g] = 1.f / (1.f + fast_exp2(-1.4426950408889634f * x0));
            f[4 * g + 1] = 1.f / (1.f + fast_exp2(-1.4426950408889634f * x1));
            f[4 * g + 2] = 1.f / (1.f + fast_exp2(-1.4426950408889634f * x2));
            f[4 * g + 3] = 1.f / (1.f + fast_exp2(-1.4426950408889634f * x3));
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) r[g] = cs[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            f[4 * g] *= r[g].x; f[4 * g + 1] *= r[g].y; f[4 * g + 2] *= r[g].z; f[4 * g + 3] *= r[g].w;
        }
        if (!first) {
#pragma unroll
            for (int g = 0; g < 8; ++g) r[g] = mix[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                f[4 * g] += r[g].x; f[4 * g + 1] += r[g].y; f[4 * g + 2] += r[g].z; f[4 * g + 3] += r[g].w;
            }
        }
        if (last) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = live ? f[j] * p.level_scale : 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                out[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4*>(c.A_buf + a_tile_off(row, c0 / 8 + g)) = pack8_u4(f + 8 * g);
        } else {
#pragma unroll
            for (int g = 0; g < 8; ++g)
                mix[static_cast<size_t>(c0 / 4 + g) * TILE_ROWS] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
        }
    }
    release_acc(c, 4, 0);
    if (last) publish_a(c);
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
    constexpr int ROWS_PER_WARP = TILE_ROWS / NW;
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

// Vocabulary projection epilogue (bias-free fc, decoders.py:90,121): fp32 logits + per-32-column-chunk
// (max, sum exp(x - max)) for beam_chunkmerge_kernel (beam.cu).
// Sparse mode: the selection (beam_chunkmerge_kernel) reads back only the `beam` <= 5 groups of 32 columns with
// the largest maxima of a row.  A thread (= row, one half of every chunk) keeps the five largest group maxima it
// has seen; a group is stored only if its maximum reaches the fifth of them -- every group of the row's final
// top five passes that test when it is produced, and ~85 % of the 52 MB of logits per step are never written.
__device__ __forceinline__ void vocab_group(WorkerCtx& c, const FusedParams& p, int chunk_idx, int colc, int grow,
                                            const uint32_t (&v)[32], float (&top)[5]) {
    const int gcol = chunk_idx * 256 + colc;
    const int valid = p.vocab - gcol;  // columns of this 32-chunk inside the vocabulary (warp-uniform)
    float cm = -INFINITY, cs = 0.f;
    if (valid >= 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
        const float cm2 = cm * 1.4426950408889634f;
#pragma unroll
        for (int j = 0; j < 32; ++j) cs += fast_exp2(fmaf(__uint_as_float(v[j]), 1.4426950408889634f, -cm2));
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
                staged_store64<true>(c.stage, c.lane, o, gbase, static_cast<size_t>(p.ld_logits) * 4, c.rows_valid_warp);
            }
        }
    }
}

__device__ __forceinline__ void epilogue_vocab_chunk(WorkerCtx& c, const FusedParams& p, int chunk_idx, float (&top)[5]) {
    const int b = acquire_acc(c, 2);
    const int row = c.quad * 32 + c.lane;
    const int grow = c.r0 + row;
    const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.quad * 32) << 16) + b * 256 + c.half * 128;
    uint32_t vv[2][32];
    tmem_ld_32x32b_x32(taddr, vv[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // group i + 1 is in flight while group i is reduced and stored
        const int colc = (c.half * 4 + i) * 32;
        tmem_ld_wait_on(vv[i & 1]);
        if (i + 1 < 4) tmem_ld_32x32b_x32(taddr + (i + 1) * 32, vv[(i + 1) & 1]);
        vocab_group(c, p, chunk_idx, colc, grow, vv[i & 1], top);
    }
    release_acc(c, 2, b);
}

// The chain kernel.  Capped at 128 registers per thread (setmaxnreg then moves them: 40 for the control warps, 168 for
// the epilogue warps): a chain CTA then takes 3/4 of the SM's register file instead of all of it, so it can start on an
// SM that still hosts a few small CTAs of other streams' kernels, and vice versa.
// (nvcc rejects __launch_bounds__ next to __maxnreg__ unless their arguments depend on a template parameter.)
// TRACE (cap_debug_fused_trace): a separate instantiation whose control warps account their waiting time per CTA into
// trace[tile * 64 + k] (SM clock cycles): issuer k = 0 total, 1 waiting for weight stages (b_full), 2 for free
// accumulators (epilogues), 3 for the A tile (a_ready / a_load), 4 for streamed A blocks; producer k = 8 total, 9 waiting
// for free ring stages (b_empty), 10 for the hidden tile / free A slots.  The production instantiation carries none of it.
template <bool PAIR, bool TRACE = false>
__global__ void __launch_bounds__(PAIR ? CHAIN_THREADS : CHAIN_THREADS, 1) __maxnreg__(PAIR ? CHAIN_MAXNREG : CHAIN_MAXNREG)
decode_chain_kernel(const __grid_constant__ FusedParams p) {
    constexpr int NBX = PAIR ? NB_PAIR : NB;
    constexpr int TSTEP = PAIR ? TILE_STEP_PAIR : 1;
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
    uint64_t* a_ready = acc_empty + 2;   // workers -> issuer: the resident A tile is complete
    uint64_t* h_ready = a_ready + 1;     // workers -> producer: the hidden tile is complete (global, proxy-fenced)
    uint64_t* a_load = h_ready + 1;      // TMA -> issuer: a tile of attention output has landed in the resident A buffer
    uint64_t* a_free = a_load + 1;       // issuer -> loader: every MMA that reads the resident A buffer has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int n_jobs = p.n_jobs;
    if (threadIdx.x == 0) flight_mark(FK_CHAIN, 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_w512)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_w2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_vocab)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_h)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            constexpr int consumers = PAIR ? 2 * NW : NW;   // pair: the leader's barriers hear both CTAs' workers
            for (int s = 0; s < NBX; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            for (int s = 0; s < A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], consumers); }
            mbar_init(a_ready, consumers);
            mbar_init(h_ready, NW);
            mbar_init(a_load, 1);
            mbar_init(a_free, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (PAIR) tmem_alloc_2sm<512>(tmem_slot); else tmem_alloc<512>(tmem_slot);
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync(); else __syncthreads();   // pair: the peer's barriers exist before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) flight_mark(FK_CHAIN_READY, 0);

    if (warp < FIRST_WORKER_WARP) {
    // control warpgroup: hand registers to the workers (the role split must sit INSIDE this branch so that
    // the register limit of each region is unambiguous to ptxas)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");   // 128 x 40 + 256 x 168 = 48 k of the 64 k registers
    if (warp == 0) {
        // ------------------------------------------------------------------ producer (weights never wait)
        // The whole warp runs the loop (uniform control flow); one elected lane issues the copies.
        uint32_t bcount = 0, acount = 0, hphase = 0;
        uint32_t tr_t0 = 0, tr_b = 0, tr_a = 0;
        if constexpr (TRACE) tr_t0 = static_cast<uint32_t>(clock64());
#define TRACED_WAIT(acc, stmt)                                                \
    do {                                                                      \
        if constexpr (TRACE) {                                                \
            const uint32_t t_ = static_cast<uint32_t>(clock64());             \
            stmt;                                                             \
            acc += static_cast<uint32_t>(clock64()) - t_;                     \
        } else {                                                              \
            stmt;                                                             \
        }                                                                     \
    } while (0)
        const uint64_t keep_policy = l2_policy_evict_last();   // weights: shared by every tile of every batch in flight
        for (int ji = 0; ji < n_jobs; ++ji) {
            const JobDesc& job = p.jobs[ji];
            const CUtensorMap* map = job.wmap == 0 ? &p.map_w512 : job.wmap == 1 ? &p.map_w2 : job.wmap == 2 ? &p.map_vocab : &p.map_wvis;
            const int chunk = job.chunk, kblocks = job.kblocks;   // read once (see epilogue_store)
            const int nch = job.ntiles / chunk;
            // pair: this CTA stages rows [rank * 128, rank * 128 + 128) of every 256-row tile pair
            const int row_base = job.row0 + (PAIR ? static_cast<int>(rank) * 128 : 0);
            const bool stream = job.a_src == JA_STREAM_HIDDEN || job.a_src == JA_STREAM_FEATURES;
            const bool from_hidden = job.a_src == JA_STREAM_HIDDEN;
            const CUtensorMap* amap = from_hidden ? &p.map_h : &p.map_feat;
            for (int c = 0; c < nch; ++c) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    for (int j = 0; j < chunk; j += TSTEP) {
                        const uint32_t s = bcount % NBX;
                        TRACED_WAIT(tr_b, mbar_wait(&b_empty[s], ((bcount / NBX) & 1) ^ 1));
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                // both CTAs signal the LEADER's barrier, which expects the two halves of the tile
                                if (leader) mbar_arrive_expect_tx(&b_full[s], 2 * BST);
                                tma_load_2d_2sm_hint(B_ring + s * BST, map, &b_full[s], kb * BLOCK_K, row_base + (c * chunk + j) * 128,
                                                     keep_policy);
                            } else {
                                mbar_arrive_expect_tx(&b_full[s], B_STAGE_BYTES);
                                tma_load_2d_hint(B_ring + s * B_STAGE_BYTES, map, &b_full[s], kb * BLOCK_K,
                                                 row_base + (c * chunk + j) * 128, keep_policy);
                            }
                        }
                        __syncwarp();
                        ++bcount;
                    }
                    if (stream) {
                        if (kb == 0 && from_hidden) {  // the workers have written (and proxy-fenced) the whole hidden tile
                            TRACED_WAIT(tr_a, mbar_wait(h_ready, hphase));
                            hphase ^= 1;
                        }
                        const uint32_t slot = acount % A_SLOTS;
                        TRACED_WAIT(tr_a, mbar_wait(&a_empty[slot], ((acount / A_SLOTS) & 1) ^ 1));
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                if (leader) mbar_arrive_expect_tx(&a_full[slot], 2 * A_KB_BYTES);
                                tma_load_2d_2sm(A_buf + slot * A_KB_BYTES, amap, &a_full[slot], kb * BLOCK_K, tile * TILE_ROWS);
                            } else {
                                mbar_arrive_expect_tx(&a_full[slot], A_KB_BYTES);
                                tma_load_2d(A_buf + slot * A_KB_BYTES, amap, &a_full[slot], kb * BLOCK_K, tile * TILE_ROWS);
                            }
                        }
                        __syncwarp();
                        ++acount;
                    }
                }
            }
        }
        if constexpr (TRACE) {
            if (p.trace != nullptr && lane == 0) {
                p.trace[static_cast<size_t>(tile) * 64 + 8] = static_cast<uint32_t>(clock64()) - tr_t0;
                p.trace[static_cast<size_t>(tile) * 64 + 9] = tr_b;
                p.trace[static_cast<size_t>(tile) * 64 + 10] = tr_a;
            }
        }
        pdl_launch_dependents();
    } else if (warp == 1 && leader) {
        // ------------------------------------------------------------------ MMA issuer (pair: the leader's only)
        // Uniform control flow for the whole warp; the elected lane issues tcgen05.mma / tcgen05.commit.
        constexpr uint32_t idesc = make_instr_desc(PAIR ? 256 : 128, PAIR ? 256 : 128);
        auto wait_consumers_at = [](uint64_t* bar, uint32_t parity, int line) {
            if constexpr (PAIR) mbar_wait_cluster_impl(bar, parity, line); else mbar_wait_impl(bar, parity, line);
        };
#define wait_consumers(bar, parity) wait_consumers_at(bar, parity, __LINE__)
        auto commit = [](uint64_t* bar) {
            if constexpr (PAIR) umma_commit_2sm(bar); else umma_commit(bar);
        };
        uint32_t bcount = 0, acount = 0, use0 = 0, use1 = 0, ar = 0, al = 0;
        int toggle = 0;
        uint32_t tr_t0 = 0, tr_b = 0, tr_acc = 0, tr_a = 0, tr_s = 0;
        if constexpr (TRACE) tr_t0 = static_cast<uint32_t>(clock64());
        for (int ji = 0; ji < n_jobs; ++ji) {
            const JobDesc& job = p.jobs[ji];
            const int chunk = job.chunk, kblocks = job.kblocks;   // read once (see epilogue_store)
            const int nch = job.ntiles / chunk;
            const bool stream = job.a_src == JA_STREAM_HIDDEN || job.a_src == JA_STREAM_FEATURES;
            if (job.a_src == JA_TMA_TILE) {
                TRACED_WAIT(tr_a, mbar_wait(a_load, al & 1));   // this job's A tile arrives by TMA (warp 2)
                ++al;
            } else if (job.flags & JF_WAIT_A) {
                TRACED_WAIT(tr_a, wait_consumers(a_ready, ar & 1));
                ++ar;
            }
            tcgen05_fence_after();
            for (int c = 0; c < nch; ++c) {
                int b = 0;
                uint32_t colbase = 0;
                if (chunk == 4) {
                    TRACED_WAIT(tr_acc, wait_consumers(&acc_empty[0], (use0 & 1) ^ 1));
                    TRACED_WAIT(tr_acc, wait_consumers(&acc_empty[1], (use1 & 1) ^ 1));
                } else {
                    b = toggle;
                    TRACED_WAIT(tr_acc, wait_consumers(&acc_empty[b], ((b ? use1 : use0) & 1) ^ 1));
                    colbase = b * 256;
                }
                tcgen05_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    const uint8_t* a_tile;
                    uint32_t slot = 0;
                    if (stream) {
                        slot = acount % A_SLOTS;
                        TRACED_WAIT(tr_s, mbar_wait(&a_full[slot], (acount / A_SLOTS) & 1));
                        a_tile = A_buf + slot * A_KB_BYTES;
                    } else {
                        a_tile = A_buf + kb * A_KB_BYTES;
                    }
                    const uint64_t a_desc = make_smem_desc(a_tile);
                    for (int j = 0; j < chunk; j += TSTEP) {
                        const uint32_t s = bcount % NBX;
                        TRACED_WAIT(tr_b, mbar_wait(&b_full[s], (bcount / NBX) & 1));
                        tcgen05_fence_after();
                        const uint64_t b_desc = make_smem_desc(B_ring + s * BST);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                if constexpr (PAIR)
                                    umma_bf16_2sm(tmem_base + colbase + j * 128, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                                  (kb | k) != 0 ? 1u : 0u);
                                else
                                    umma_bf16(tmem_base + colbase + j * 128, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                              (kb | k) != 0 ? 1u : 0u);
                            }
                            commit(&b_empty[s]);  // stage reusable (in both CTAs) once these MMAs have read it
                        }
                        __syncwarp();
                        ++bcount;
                    }
                    if (stream) {
                        if (elect_one_sync()) commit(&a_empty[slot]);
                        __syncwarp();
                        ++acount;
                    }
                }
                if (elect_one_sync()) {
                    if (chunk == 4) {
                        commit(&acc_full[0]);
                        commit(&acc_full[1]);
                    } else {
                        commit(&acc_full[b]);
                    }
                }
                __syncwarp();
                if (chunk == 4) {
                    ++use0; ++use1;
                    toggle = 0;
                } else {
                    if (b) ++use1; else ++use0;
                    toggle ^= 1;
                }
            }
            // the next job's A tile comes by TMA into the resident buffer: free once every MMA issued so far has read it
            if (ji + 1 < n_jobs && p.jobs[ji + 1].a_src == JA_TMA_TILE) {
                if (elect_one_sync()) commit(a_free);
                __syncwarp();
            }
        }
#undef wait_consumers
        if constexpr (TRACE) {
            if (p.trace != nullptr && lane == 0) {
                unsigned long long* tr = p.trace + static_cast<size_t>(tile) * 64;
                tr[0] = static_cast<uint32_t>(clock64()) - tr_t0;
                tr[1] = tr_b; tr[2] = tr_acc; tr[3] = tr_a; tr[4] = tr_s;
            }
        }
#undef TRACED_WAIT
        pdl_launch_dependents();
    } else if (warp == 2) {
        // ------------------------------------------------------------------ A-tile loader
        uint32_t loads = 0, frees = 0;
        for (int ji = 0; ji < n_jobs; ++ji) {
            const JobDesc& job = p.jobs[ji];
            if (job.a_src != JA_TMA_TILE) continue;
            if (loads == 0) pdl_wait();  // the attention kernel that wrote the tile is a stream predecessor
            if (ji > 0) {                // the buffer is still the operand of earlier jobs
                mbar_wait(a_free, frees & 1);
                ++frees;
            }
            if (elect_one_sync()) {
                if constexpr (PAIR) {
                    if (leader) mbar_arrive_expect_tx(a_load, 2 * A_SLOTS * A_KB_BYTES);
                    for (int kb = 0; kb < A_SLOTS; ++kb)
                        tma_load_2d_2sm(A_buf + kb * A_KB_BYTES, &p.map_att, a_load, kb * BLOCK_K, job.att_row0 + tile * TILE_ROWS);
                } else {
                    mbar_arrive_expect_tx(a_load, A_SLOTS * A_KB_BYTES);
                    for (int kb = 0; kb < A_SLOTS; ++kb)
                        tma_load_2d(A_buf + kb * A_KB_BYTES, &p.map_att, a_load, kb * BLOCK_K, job.att_row0 + tile * TILE_ROWS);
                }
            }
            __syncwarp();
            ++loads;
        }
    }
    } else {
        // ------------------------------------------------------------------ workers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
        WorkerCtx c;
        c.A_buf = A_buf;
        c.ww = warp - FIRST_WORKER_WARP;
        c.quad = warp & 3;
        c.half = c.ww >> 2;
        c.lane = lane;
        c.wtid = threadIdx.x - FIRST_WORKER_WARP * 32;
        c.stage = smem + OFF_STAGE + c.ww * 32 * STAGE_PITCH;
        c.s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
        c.s_cbias = reinterpret_cast<float*>(smem + OFF_CBIAS);
        c.s_gamma = reinterpret_cast<float*>(smem + OFF_GAMMA);
        c.s_beta = reinterpret_cast<float*>(smem + OFF_BETA);
        c.s_stat = reinterpret_cast<float*>(smem + OFF_STAT);
        c.acc_full = acc_full;
        c.acc_empty = acc_empty;
        c.a_ready = a_ready;
        c.h_ready = h_ready;
        c.pair_peer = PAIR && !leader;
        c.tmem_base = tmem_base;
        c.use0 = c.use1 = 0;
        c.toggle = 0;
        c.tile = tile;
        c.r0 = tile * TILE_ROWS;
        c.rows_valid_warp = max(0, min(32, p.R - (c.r0 + c.quad * 32)));

        float bias_reg = 0.f;
        pdl_wait();  // everything this chain reads was written by stream predecessors
        if (p.start_embed) {
            embed_phase(c, p);
            workers_sync();   // the residual tile was written warp-per-row, the epilogues read it thread-per-row
            publish_a(c);
        }
        float* mix_stream = nullptr;   // the meshed mix accumulates in the stream the first level's gate job names
        for (int ji = 0; ji < n_jobs; ++ji) {
            const JobDesc& job = p.jobs[ji];
            if (job.epi == JE_STORE) {
                const int nch = job.ntiles / 2;
                for (int ch = 0; ch < nch; ++ch) epilogue_store(c, p, job, ch, nch, bias_reg);
                if (job.flags & JF_HIDDEN_DONE) {
                    fence_proxy_async();  // hidden tile (global, generic proxy) -> TMA reads (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(h_ready);
                }
            } else if (job.epi == JE_LN) {
                epilogue_layernorm(c, p, job);
            } else if (job.epi == JE_F32) {
                const int nch = job.ntiles / 2;
                for (int ch = 0; ch < nch; ++ch) epilogue_f32(c, p, ch);
            } else if (job.epi == JE_ALPHA) {
                if (job.flags & JF_FIRST_LEVEL) mix_stream = job.res_out;
                epilogue_alpha(c, p, job, mix_stream);
            } else {   // JE_VOCAB: always the last job of its chain
                pdl_launch_dependents();
                const int vchunks = job.ntiles / 2;
                float top[5] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
                for (int ch = 0; ch < vchunks; ++ch) epilogue_vocab_chunk(c, p, ch, top);
            }
        }
        pdl_launch_dependents();
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync(); else __syncthreads();   // pair: the leader's MMAs read the peer's shared memory
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
    }
    if (threadIdx.x == 0) {
        flight_mark(FK_CHAIN, 1);
        flight_mark(FK_CHAIN_READY, 1);
    }
}

int launch_chain(FusedParams& p, int tiles, bool use_pairs, cudaStream_t stream);
int set_chain_smem_attributes();

}  // namespace

// ------------------------------------------------------------------------------------------ host side
// Stacked bf16 copies of a decoder's projection weights, so that one TMA tensor map serves every GEMM of a chain:
//   K = 512 stack, per layer (plain decoder, 5120 rows):  q|k|v 1536, self fc_o 512, cross fc_q 512, cross fc_o 512, fc1 2048
//                  per layer (meshed decoder, 8192 rows): q|k|v 1536, self fc_o 512, cross fc_q 512, gates' s-part 3 x 512,
//                                                          cross fc_o 512, gates' c-part 3 x 512, fc1 2048
//   fc2 stack [layers * 512][2048].
// The meshed gates fc_alphas.i are [512][1024] over [s ; c_i] (decoders.py:60-62): W_i[:, :512] multiplies the
// self-attention output s (one N = 1536 GEMM for the three levels while s is the resident tile), W_i[:, 512:] the
// level's cross-attention output c_i.
// One set can serve any number of handles (cap_fused_desc::stacked): engines that pipeline independent batches then
// stream the SAME addresses, which the evict_last hint keeps L2-resident.
struct cap_fused_weights {
    void* w512 = nullptr;
    void* w2 = nullptr;
    int n_layers = 0;
    int levels = 0;          // 0: plain decoder
    int rows_per_layer = 0;
    int off_o1 = 0, off_q = 0, off_gate_s = 0, off_o2 = 0, off_gate_c = 0, off_w1 = 0;
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
    const int levels = layers[0].n_levels;
    CAP_REQUIRE(levels == 0 || levels == MAX_LEVELS, "cap_fused_weights_create: the meshed chains are written for %d encoder levels", MAX_LEVELS);
    cap_fused_weights* f = new cap_fused_weights();
    f->n_layers = n_layers;
    f->levels = levels;
    f->off_o1 = 3 * FD;
    f->off_q = f->off_o1 + FD;
    f->off_gate_s = f->off_q + FD;
    f->off_o2 = f->off_gate_s + levels * FD;
    f->off_gate_c = f->off_o2 + FD;
    f->off_w1 = f->off_gate_c + levels * FD;
    f->rows_per_layer = f->off_w1 + FDFF;
    auto fail = [&](int rc) { cap_fused_weights_destroy(f); return rc; };
    const size_t w512_elems = static_cast<size_t>(n_layers) * f->rows_per_layer * FD;
    const size_t w2_elems = static_cast<size_t>(n_layers) * FD * FDFF;
    if (cudaMalloc(&f->w512, w512_elems * 2) != cudaSuccess || cudaMalloc(&f->w2, w2_elems * 2) != cudaSuccess)
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: cudaMalloc of the stacked weights failed"));
    for (int l = 0; l < n_layers; ++l) {
        const cap_fused_layer& w = layers[l];
        CAP_REQUIRE(w.n_levels == levels, "cap_fused_weights_create: layers disagree on the number of encoder levels");
        bf16* base = static_cast<bf16*>(f->w512) + static_cast<size_t>(l) * f->rows_per_layer * FD;
        const void* srcs[5] = {w.w_qkv, w.w_o1, w.w_q, w.w_o2, w.w_fc1};
        const int offs[5] = {0, f->off_o1, f->off_q, f->off_o2, f->off_w1};
        const int rows[5] = {3 * FD, FD, FD, FD, FDFF};
        for (int i = 0; i < 5; ++i) {
            if (!srcs[i]) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_weights_create: null weight"));
            if (cudaMemcpy(base + static_cast<size_t>(offs[i]) * FD, srcs[i], static_cast<size_t>(rows[i]) * FD * 2,
                           cudaMemcpyDeviceToDevice) != cudaSuccess)
                return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: weight copy failed"));
        }
        for (int i = 0; i < levels; ++i) {   // fc_alphas.i [512][1024] -> s-part and c-part, [512][512] each
            const bf16* src = static_cast<const bf16*>(w.w_alpha[i]);
            if (!src) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_weights_create: null gate weight"));
            bf16* dst_s = base + static_cast<size_t>(f->off_gate_s + i * FD) * FD;
            bf16* dst_c = base + static_cast<size_t>(f->off_gate_c + i * FD) * FD;
            if (cudaMemcpy2D(dst_s, FD * 2, src, 2 * FD * 2, FD * 2, FD, cudaMemcpyDeviceToDevice) != cudaSuccess ||
                cudaMemcpy2D(dst_c, FD * 2, src + FD, 2 * FD * 2, FD * 2, FD, cudaMemcpyDeviceToDevice) != cudaSuccess)
                return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_weights_create: gate weight copy failed"));
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
    cap_fused_layer layers[MAX_FUSED_LAYERS];   // bias / LayerNorm pointers (the weight pointers are not used after creation)
    cap_fused_weights* stacked = nullptr;   // the stacked weights the tensor maps point into
    bool owns_stacked = false;              // false: cap_fused_desc::stacked, owned by the caller
    int n_layers = 0, levels = 0, max_rows = 0;
    int tiles = 0;
    bf16* qkv_cache = nullptr;   // [layers][T][R][1536]
    bf16* q_out = nullptr;       // [R][512] cross-attention queries
    float *res_c = nullptr, *res_mix = nullptr;   // meshed side streams
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
    CAP_REQUIRE(d->layers != nullptr && d->att_in != nullptr && d->q_out != nullptr, "cap_fused_create: null layers / att_in / q_out");
    const int levels = d->layers[0].n_levels;
    CAP_REQUIRE(levels == 0 || levels == MAX_LEVELS, "cap_fused_create: the meshed chains are written for %d encoder levels", MAX_LEVELS);
    CAP_REQUIRE(d->stacked == nullptr || (d->stacked->n_layers == d->n_layers && d->stacked->levels == levels),
                "cap_fused_create: stacked weights of another model");
    CAP_PROPAGATE(install_fault_buffer());
    cap_install_flight();
    cap_fused_decoder* f = new cap_fused_decoder();
    FusedParams& p = f->base;
    memset(&p, 0, sizeof(p));
    const int L = d->n_layers;
    f->n_layers = L;
    f->levels = levels;
    f->max_rows = d->max_rows;
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
        f->layers[l] = w;
        const float* need[12] = {w.b_qkv, w.b_o1, w.ln1_g, w.ln1_b, w.b_q, w.b_o2, w.ln2_g, w.ln2_b, w.b_fc1, w.b_fc2, w.ln3_g, w.ln3_b};
        for (const float* q : need)
            if (!q) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_create: null bias / LayerNorm parameter"));
        for (int i = 0; i < levels; ++i)
            if (!w.b_alpha[i]) return fail(cap_set_error(CAP_ERR_INVALID, "cap_fused_create: null gate bias"));
    }
    void* const w512 = f->stacked->w512;
    void* const w2 = f->stacked->w2;
    const int rpl = f->stacked->rows_per_layer;
    int rc = cap_gemm::make_tmap(&p.map_w512, w512, L * rpl, FD, FD, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_w2, w2, L * FD, FDFF, FDFF, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_vocab, d->w_vocab, d->vocab, FD, FD, 128);
    if (rc != CAP_OK) return fail(rc);
    p.tokens = d->tokens; p.word_emb = static_cast<const bf16*>(d->word_emb); p.word_pos = d->word_pos; p.pad_idx = d->pad_idx;
    p.padflag = d->padflag;
    f->qkv_cache = static_cast<bf16*>(d->qkv_cache);
    p.logits = d->logits; p.ld_logits = d->ld_logits; p.part_ms = d->part_ms;
    p.vocab = d->vocab;
    p.vocab_tiles = ((d->vocab + 255) / 256) * 2;
    p.stat_chunks = ((d->vocab + 255) / 256) * 8;
    p.T = d->max_len; p.beam = d->beam;
    p.level_scale = levels > 0 ? 1.0f / sqrtf(static_cast<float>(levels)) : 1.0f;
    p.aux_cols = levels * FD;
    const size_t tiles = f->tiles;
    const size_t stream_bytes = tiles * TILE_ROWS * FD * 4;
    void *res = nullptr, *hb = nullptr, *aux = nullptr, *rc_ = nullptr, *rm = nullptr;
    bool ok = cudaMalloc(&res, stream_bytes) == cudaSuccess && cudaMalloc(&hb, tiles * TILE_ROWS * FDFF * 2) == cudaSuccess;
    if (ok && levels > 0)
        ok = cudaMalloc(&aux, stream_bytes * levels) == cudaSuccess && cudaMalloc(&rc_, stream_bytes) == cudaSuccess &&
             cudaMalloc(&rm, stream_bytes) == cudaSuccess;
    p.res = static_cast<float*>(res); p.hbuf = static_cast<bf16*>(hb); p.aux = static_cast<float*>(aux);
    f->res_c = static_cast<float*>(rc_); f->res_mix = static_cast<float*>(rm);
    if (!ok) return fail(cap_set_error(CAP_ERR_CUDA, "cap_fused_create: cudaMalloc of the scratch tiles failed"));
    cudaMemset(res, 0, stream_bytes);
    cudaMemset(hb, 0, tiles * TILE_ROWS * FDFF * 2);
    if (levels > 0) {
        cudaMemset(aux, 0, stream_bytes * levels);
        cudaMemset(rc_, 0, stream_bytes);
        cudaMemset(rm, 0, stream_bytes);
    }
    rc = cap_gemm::make_tmap(&p.map_h, hb, static_cast<int>(tiles) * TILE_ROWS, FDFF, FDFF, 128);
    if (rc != CAP_OK) return fail(rc);
    // attention outputs: [levels][max_rows][512] (one level for the plain decoder)
    rc = cap_gemm::make_tmap(&p.map_att, d->att_in, (levels > 0 ? levels : 1) * d->max_rows, FD, FD, 128);
    if (rc != CAP_OK) return fail(rc);
    f->q_out = static_cast<bf16*>(d->q_out);
    f->has_att = true;
    f->full_logits = getenv("OPENVIIC_FULL_LOGITS") && atoi(getenv("OPENVIIC_FULL_LOGITS")) != 0;
    f->use_pairs = !(getenv("OPENVIIC_CHAIN_PAIR") && atoi(getenv("OPENVIIC_CHAIN_PAIR")) == 0);
    rc = set_chain_smem_attributes();
    if (rc != CAP_OK) return fail(rc);
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
    cudaFree(f->base.hbuf);
    cudaFree(f->base.aux);
    cudaFree(f->res_c);
    cudaFree(f->res_mix);
    delete f;
    return CAP_OK;
}

static unsigned long long* g_fused_trace = nullptr;
extern "C" int cap_debug_fused_trace(unsigned long long* device_buffer) {
    g_fused_trace = device_buffer;
    return CAP_OK;
}

namespace {
int set_chain_smem_attributes() {
    static cap_device_once once;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&once.done, __ATOMIC_ACQUIRE) & bit) return CAP_OK;
    if (cudaFuncSetAttribute(decode_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(decode_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(decode_chain_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(decode_chain_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM) != cudaSuccess)
        return cap_set_error(CAP_ERR_CUDA, "chain kernels: cannot reserve %u bytes of shared memory", FUSED_SMEM);
    __atomic_fetch_or(&once.done, bit, __ATOMIC_RELEASE);
    return CAP_OK;
}
// ---- job lists ------------------------------------------------------------------------------------------------
struct JobList {
    FusedParams& p;
    const cap_fused_decoder& f;
    int layer;
    JobDesc& add() {
        if (p.n_jobs >= MAX_JOBS) p.n_jobs = MAX_JOBS - 1;   // reported by the caller's size check (never hit by the lists below)
        JobDesc& j = p.jobs[p.n_jobs++];
        memset(&j, 0, sizeof(j));
        j.kblocks = FD / BLOCK_K;
        j.chunk = 2;
        j.a_src = JA_RESIDENT;
        return j;
    }
    int base() const { return layer * f.stacked->rows_per_layer; }
    const cap_fused_layer& w() const { return f.layers[layer]; }
    // q|k|v of the new token of `layer` -> cache slot of step t
    void qkv() {
        JobDesc& j = add();
        j.row0 = base(); j.ntiles = 12; j.epi = JE_STORE; j.bias = w().b_qkv;
        j.dst = f.qkv_cache + (static_cast<size_t>(layer) * p.T + p.t) * p.R * 3 * FD;
        j.ld_dst = 3 * FD;
    }
    void vocab() {
        JobDesc& j = add();
        j.wmap = 2; j.row0 = 0; j.ntiles = p.vocab_tiles; j.epi = JE_VOCAB;
    }
    // fc_o of an attention + residual + LayerNorm; the tile of attention outputs arrives by TMA
    void attention_out(const float* bias, const float* g, const float* b, int weight_off, int att_row0, const float* res_in, float* res_out) {
        JobDesc& j = add();
        j.row0 = base() + weight_off; j.ntiles = 4; j.chunk = 4; j.a_src = JA_TMA_TILE; j.att_row0 = att_row0; j.epi = JE_LN;
        j.bias = bias; j.gamma = g; j.beta = b; j.res_in = res_in; j.res_out = res_out;
    }
    void ffn_and_next() {
        JobDesc& a = add();   // fc1 + ReLU -> hidden scratch
        a.row0 = base() + f.stacked->off_w1; a.ntiles = 16; a.epi = JE_STORE; a.flags = JF_RELU | JF_WHOLE_TILES | JF_HIDDEN_DONE;
        a.bias = w().b_fc1; a.dst = p.hbuf; a.ld_dst = FDFF;
        JobDesc& b = add();   // fc2 + residual + LayerNorm, rows fed <pad> zeroed (decoders.py:26)
        b.wmap = 1; b.row0 = layer * FD; b.ntiles = 4; b.chunk = 4; b.kblocks = FDFF / BLOCK_K; b.a_src = JA_STREAM_HIDDEN; b.epi = JE_LN;
        b.bias = w().b_fc2; b.gamma = w().ln3_g; b.beta = w().ln3_b; b.res_in = p.res; b.res_out = p.res;
        b.zero_rows = p.padflag + static_cast<size_t>(p.t) * p.R;
        if (layer + 1 < f.n_layers) { ++layer; qkv(); --layer; } else vocab();
    }
};
}  // namespace

extern "C" int cap_fused_chain(cap_fused_decoder* f, int chain, int layer, int t, int B, cap_stream_t stream) {
    CAP_REQUIRE(f != nullptr, "cap_fused_chain: null handle");
    FusedParams p = f->base;
    CAP_REQUIRE(t >= 0 && t < p.T, "cap_fused_chain: step %d outside [0,%d)", t, p.T);
    CAP_REQUIRE(chain >= CAP_CHAIN_EMBED_QKV && chain <= CAP_CHAIN_FFN && layer >= 0 && layer < f->n_layers,
                "cap_fused_chain: bad chain %d / layer %d", chain, layer);
    const int R = B * p.beam;
    const int tiles = (R + TILE_ROWS - 1) / TILE_ROWS;
    CAP_REQUIRE(B > 0 && tiles <= f->tiles && R <= f->max_rows, "cap_fused_chain: batch %d exceeds the reservation", B);
    p.t = t; p.R = R; p.B = B;
    p.sparse_logits = (f->full_logits || p.beam > 5) ? 0 : 1;
    // debug (cap_debug_fused_trace): one region of 64 tiles x 64 words per (chain kind, layer), else nullptr
    p.trace = g_fused_trace ? g_fused_trace + static_cast<size_t>(chain * MAX_FUSED_LAYERS + layer) * 64 * 64 : nullptr;
    p.start_embed = 0;
    p.n_jobs = 0;
    JobList jl{p, *f, layer};
    const cap_fused_layer& w = f->layers[layer];
    const cap_fused_weights& sw = *f->stacked;
    if (chain == CAP_CHAIN_EMBED_QKV) {          // x = Emb + pos; q|k|v of layer 0 -> cache
        CAP_REQUIRE(layer == 0, "cap_fused_chain: the embedding chain belongs to layer 0");
        p.start_embed = 1;
        jl.qkv();
    } else if (chain == CAP_CHAIN_SELF_OUT) {    // self fc_o + LN; cross fc_q -> q_out; meshed: the gates' s-part
        jl.attention_out(w.b_o1, w.ln1_g, w.ln1_b, sw.off_o1, 0, p.res, p.res);
        JobDesc& q = jl.add();
        q.row0 = jl.base() + sw.off_q; q.ntiles = 4; q.epi = JE_STORE; q.bias = w.b_q; q.dst = f->q_out; q.ld_dst = FD;
        if (f->levels > 0) {
            JobDesc& g = jl.add();
            g.row0 = jl.base() + sw.off_gate_s; g.ntiles = 4 * f->levels; g.epi = JE_F32;
        }
    } else if (f->levels == 0) {                 // cross fc_o + LN; FFN + LN; next layer's q|k|v or the vocabulary
        jl.attention_out(w.b_o2, w.ln2_g, w.ln2_b, sw.off_o2, 0, p.res, p.res);
        jl.ffn_and_next();
    } else {                                     // meshed: per level cross fc_o + LN -> c_i, gate_i, mix; then as above
        for (int i = 0; i < f->levels; ++i) {
            jl.attention_out(w.b_o2, w.ln2_g, w.ln2_b, sw.off_o2, i * f->max_rows, p.res, f->res_c);   // same enc_attn weights per level (decoders.py:35,56)
            JobDesc& a = jl.add();
            a.row0 = jl.base() + sw.off_gate_c + i * FD; a.ntiles = 4; a.chunk = 4; a.epi = JE_ALPHA; a.bias = w.b_alpha[i];
            a.aux_col0 = i * FD; a.res_in = f->res_c;
            a.flags = (i == 0 ? JF_FIRST_LEVEL : 0) | (i == f->levels - 1 ? JF_LAST_LEVEL : 0);
            a.res_out = (i == f->levels - 1) ? p.res : f->res_mix;   // the last level hands the scaled mix on as the FFN's residual
        }
        jl.ffn_and_next();
    }
    return launch_chain(p, tiles, f->use_pairs, static_cast<cudaStream_t>(stream));
}

namespace {
// No programmatic dependent launch for a chain: its CTAs each take a whole SM, and an early-started chain
// would hold them idle in griddepcontrol.wait until the attention kernel before it has drained.
// CTA pairs (OPENVIIC_CHAIN_PAIR=0: single CTAs): clusters of two adjacent tiles, the second CTA of the last
// pair is a dummy when the tile count is odd (every access of it is guarded by the row count).
int launch_chain(FusedParams& p, int tiles, bool use_pairs, cudaStream_t stream) {
    CAP_REQUIRE(p.n_jobs >= 1 && p.n_jobs <= MAX_JOBS, "chain: job list overflow");
    for (int ji = 0; ji < p.n_jobs; ++ji) {   // which jobs start from a freshly written resident tile
        JobDesc& j = p.jobs[ji];
        if (j.a_src != JA_RESIDENT) continue;
        const bool fresh = ji == 0 ? p.start_embed != 0
                                   : (p.jobs[ji - 1].epi == JE_LN || (p.jobs[ji - 1].epi == JE_ALPHA && (p.jobs[ji - 1].flags & JF_LAST_LEVEL)));
        if (fresh) j.flags |= JF_WAIT_A;
        CAP_REQUIRE(ji > 0 || p.start_embed, "chain: the first job has no A tile");
    }
    CAP_REQUIRE(p.trace == nullptr || tiles <= 64, "chain: the trace buffer holds 64 tiles");
    if (use_pairs) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((tiles + 1) / 2 * 2);
        cfg.blockDim = dim3(CHAIN_THREADS);
        cfg.dynamicSmemBytes = FUSED_SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (p.trace != nullptr) cudaLaunchKernelEx(&cfg, decode_chain_kernel<true, true>, p);
        else cudaLaunchKernelEx(&cfg, decode_chain_kernel<true>, p);
    } else if (p.trace != nullptr) {
        decode_chain_kernel<false, true><<<dim3(tiles), dim3(CHAIN_THREADS), FUSED_SMEM, stream>>>(p);
    } else {
        decode_chain_kernel<false><<<dim3(tiles), dim3(CHAIN_THREADS), FUSED_SMEM, stream>>>(p);
    }
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("decode_chain_kernel");
}
}  // namespace

// ------------------------------------------------------------------------------------------ encoder chains
// The encoder (encoders.py:17-40, vision_embeddings.py:15-20) on the same chain kernel: a tile of 128 visual-token rows
// is carried through
//   stage 0     : vision projection (K = d_feature, the raw features streamed as the A operand) + LayerNorm + position
//                 table, then layer 0's q|k|v projection;
//   stage 1 + l : layer l's fc_o + residual + LayerNorm, fc1, fc2 + residual + LayerNorm (padded rows zeroed, the level
//                 output also stored row-major), then layer l + 1's q|k|v -- or, from the level outputs the decoder
//                 attends to, the decoder layers' cross-attention K|V projections;
// with the encoder's self-attention kernel (attention.cu) between the stages: 1 + layers chain launches instead of
// ~7 GEMM launches per layer, and no activation except q|k|v, the hidden tile and the final K|V leaves the SM.
struct cap_enc_chain_weights {
    void* w512 = nullptr;   // per layer: q|k|v 1536, fc_o 512, fc1 2048 rows; then n_kv x 1024 rows of cross K|V projections
    void* w2 = nullptr;     // [layers * 512][2048]
    int n_layers = 0, n_kv = 0;
};

extern "C" int cap_enc_chain_weights_destroy(cap_enc_chain_weights* w) {
    if (!w) return CAP_OK;
    cudaFree(w->w512);
    cudaFree(w->w2);
    delete w;
    return CAP_OK;
}

namespace {
constexpr int ENC_ROWS_PER_LAYER = 3 * FD + FD + FDFF;   // q|k|v, fc_o, fc1
constexpr int MAX_ENC_LAYERS = 6;
}  // namespace

extern "C" int cap_enc_chain_weights_create(const cap_enc_chain_desc* d, cap_enc_chain_weights** out) {
    CAP_REQUIRE(d && out && d->layers, "cap_enc_chain_weights_create: null pointer");
    CAP_REQUIRE(d->n_layers >= 1 && d->n_layers <= MAX_ENC_LAYERS && d->n_kv >= 0 && d->n_kv <= CAP_ENC_MAX_KV,
                "cap_enc_chain_weights_create: bad layer / projection count");
    cap_enc_chain_weights* f = new cap_enc_chain_weights();
    f->n_layers = d->n_layers;
    f->n_kv = d->n_kv;
    auto fail = [&](int rc) { cap_enc_chain_weights_destroy(f); return rc; };
    const size_t rows512 = static_cast<size_t>(d->n_layers) * ENC_ROWS_PER_LAYER + static_cast<size_t>(d->n_kv) * 2 * FD;
    if (cudaMalloc(&f->w512, rows512 * FD * 2) != cudaSuccess ||
        cudaMalloc(&f->w2, static_cast<size_t>(d->n_layers) * FD * FDFF * 2) != cudaSuccess)
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_enc_chain_weights_create: cudaMalloc of the stacked weights failed"));
    auto copy = [&](bf16* dst, const void* src, size_t rows, size_t cols) {
        return src != nullptr && cudaMemcpy(dst, src, rows * cols * 2, cudaMemcpyDeviceToDevice) == cudaSuccess;
    };
    bf16* w512 = static_cast<bf16*>(f->w512);
    for (int l = 0; l < d->n_layers; ++l) {
        const cap_enc_layer& w = d->layers[l];
        bf16* base = w512 + static_cast<size_t>(l) * ENC_ROWS_PER_LAYER * FD;
        if (!copy(base, w.w_qkv, 3 * FD, FD) || !copy(base + static_cast<size_t>(3 * FD) * FD, w.w_o, FD, FD) ||
            !copy(base + static_cast<size_t>(4 * FD) * FD, w.w_fc1, FDFF, FD) ||
            !copy(static_cast<bf16*>(f->w2) + static_cast<size_t>(l) * FD * FDFF, w.w_fc2, FD, FDFF))
            return fail(cap_set_error(CAP_ERR_INVALID, "cap_enc_chain_weights_create: null weight or copy failed (layer %d)", l));
    }
    for (int i = 0; i < d->n_kv; ++i)
        if (!copy(w512 + (static_cast<size_t>(d->n_layers) * ENC_ROWS_PER_LAYER + static_cast<size_t>(i) * 2 * FD) * FD, d->w_kv[i], 2 * FD, FD))
            return fail(cap_set_error(CAP_ERR_INVALID, "cap_enc_chain_weights_create: null K|V weight or copy failed (%d)", i));
    *out = f;
    return CAP_OK;
}

struct cap_enc_chains {
    FusedParams base;
    cap_enc_chain_desc desc;
    cap_enc_layer layers[MAX_ENC_LAYERS];
    cap_enc_chain_weights* stacked = nullptr;
    bool owns_stacked = false;
    int tiles = 0;
    bool use_pairs = true;
};

extern "C" int cap_enc_chains_destroy(cap_enc_chains* f) {
    if (!f) return CAP_OK;
    if (f->owns_stacked) cap_enc_chain_weights_destroy(f->stacked);
    cudaFree(f->base.res);
    cudaFree(f->base.hbuf);
    delete f;
    return CAP_OK;
}

extern "C" int cap_enc_chains_create(const cap_enc_chain_desc* d, cap_enc_chains** out) {
    CAP_REQUIRE(d && out && d->layers, "cap_enc_chains_create: null pointer");
    CAP_REQUIRE(d->d_model == FD && d->d_ff == FDFF, "cap_enc_chains_create: needs d_model 512, d_ff 2048");
    CAP_REQUIRE(d->d_feature >= BLOCK_K && d->d_feature % BLOCK_K == 0, "cap_enc_chains_create: d_feature must be a multiple of %d", BLOCK_K);
    CAP_REQUIRE(d->n_layers >= 1 && d->n_layers <= MAX_ENC_LAYERS && d->n_kv >= 0 && d->n_kv <= CAP_ENC_MAX_KV, "cap_enc_chains_create: bad counts");
    CAP_REQUIRE(d->max_rows > 0 && d->w_vis && d->b_vis && d->ln0_g && d->ln0_b && d->pos && d->feats && d->qkv_out && d->att_in && d->levels_out,
                "cap_enc_chains_create: null tensor");
    CAP_REQUIRE(d->stacked == nullptr || (d->stacked->n_layers == d->n_layers && d->stacked->n_kv == d->n_kv),
                "cap_enc_chains_create: stacked weights of another model");
    CAP_PROPAGATE(install_fault_buffer());
    cap_install_flight();
    cap_enc_chains* f = new cap_enc_chains();
    FusedParams& p = f->base;
    memset(&p, 0, sizeof(p));
    f->desc = *d;
    for (int l = 0; l < d->n_layers; ++l) f->layers[l] = d->layers[l];
    f->desc.layers = f->layers;
    f->tiles = ((d->max_rows + TILE_ROWS - 1) / TILE_ROWS + 1) / 2 * 2;
    auto fail = [&](int rc) { cap_enc_chains_destroy(f); return rc; };
    if (d->stacked) {
        f->stacked = const_cast<cap_enc_chain_weights*>(d->stacked);
    } else {
        const int rc0 = cap_enc_chain_weights_create(d, &f->stacked);
        if (rc0 != CAP_OK) return fail(rc0);
        f->owns_stacked = true;
    }
    const int rows512 = d->n_layers * ENC_ROWS_PER_LAYER + d->n_kv * 2 * FD;
    int rc = cap_gemm::make_tmap(&p.map_w512, f->stacked->w512, rows512, FD, FD, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_w2, f->stacked->w2, d->n_layers * FD, FDFF, FDFF, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_wvis, d->w_vis, FD, d->d_feature, d->d_feature, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_feat, d->feats, d->max_rows, d->d_feature, d->d_feature, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_att, d->att_in, d->max_rows, FD, FD, 128);
    if (rc == CAP_OK) rc = cap_gemm::make_tmap(&p.map_vocab, f->stacked->w512, rows512, FD, FD, 128);     // unused by encoder jobs: any valid map
    if (rc != CAP_OK) return fail(rc);
    const size_t tiles = f->tiles;
    void *res = nullptr, *hb = nullptr;
    if (cudaMalloc(&res, tiles * TILE_ROWS * FD * 4) != cudaSuccess || cudaMalloc(&hb, tiles * TILE_ROWS * FDFF * 2) != cudaSuccess) {
        cudaFree(res);
        return fail(cap_set_error(CAP_ERR_CUDA, "cap_enc_chains_create: cudaMalloc of the scratch tiles failed"));
    }
    p.res = static_cast<float*>(res);
    p.hbuf = static_cast<bf16*>(hb);
    cudaMemset(res, 0, tiles * TILE_ROWS * FD * 4);
    cudaMemset(hb, 0, tiles * TILE_ROWS * FDFF * 2);
    rc = cap_gemm::make_tmap(&p.map_h, hb, static_cast<int>(tiles) * TILE_ROWS, FDFF, FDFF, 128);
    if (rc != CAP_OK) return fail(rc);
    p.level_scale = 1.f;
    p.beam = 1;
    f->use_pairs = !(getenv("OPENVIIC_CHAIN_PAIR") && atoi(getenv("OPENVIIC_CHAIN_PAIR")) == 0);
    rc = set_chain_smem_attributes();
    if (rc != CAP_OK) return fail(rc);
    *out = f;
    return CAP_OK;
}

// stage 0: vision projection + LayerNorm + positions, q|k|v of layer 0; stage 1 + l: the rest of layer l (see above)
extern "C" int cap_enc_chain(cap_enc_chains* f, int stage, int rows, int n_tokens, cap_stream_t stream) {
    CAP_REQUIRE(f != nullptr, "cap_enc_chain: null handle");
    const cap_enc_chain_desc& d = f->desc;
    CAP_REQUIRE(stage >= 0 && stage <= d.n_layers, "cap_enc_chain: stage %d outside [0,%d]", stage, d.n_layers);
    CAP_REQUIRE(rows > 0 && rows <= d.max_rows && n_tokens > 0, "cap_enc_chain: %d rows exceed the reservation", rows);
    FusedParams p = f->base;
    const int tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    p.R = rows; p.B = rows; p.t = 0; p.T = 1;
    p.pos_rows = n_tokens;
    p.start_embed = 0;
    p.n_jobs = 0;
    p.trace = g_fused_trace ? g_fused_trace + static_cast<size_t>(stage % (3 * MAX_FUSED_LAYERS)) * 64 * 64 : nullptr;
    auto add = [&]() -> JobDesc& {
        JobDesc& j = p.jobs[p.n_jobs < MAX_JOBS ? p.n_jobs++ : MAX_JOBS - 1];
        memset(&j, 0, sizeof(j));
        j.kblocks = FD / BLOCK_K;
        j.chunk = 2;
        j.a_src = JA_RESIDENT;
        return j;
    };
    auto qkv = [&](int layer) {
        JobDesc& j = add();
        j.row0 = layer * ENC_ROWS_PER_LAYER; j.ntiles = 12; j.epi = JE_STORE; j.bias = f->layers[layer].b_qkv;
        j.dst = static_cast<bf16*>(d.qkv_out); j.ld_dst = 3 * FD;
    };
    if (stage == 0) {
        JobDesc& v = add();   // x = LN(W_vis . features + b) + pos   (vision_embeddings.py:18, encoders.py:36)
        v.wmap = 3; v.row0 = 0; v.ntiles = 4; v.chunk = 4; v.kblocks = d.d_feature / BLOCK_K; v.a_src = JA_STREAM_FEATURES;
        v.epi = JE_LN; v.flags = JF_NO_RESIDUAL; v.bias = d.b_vis; v.gamma = d.ln0_g; v.beta = d.ln0_b; v.pos = d.pos;
        v.res_in = p.res; v.res_out = p.res;
        qkv(0);
    } else {
        const int l = stage - 1;
        const cap_enc_layer& w = f->layers[l];
        JobDesc& o = add();   // LN(x + fc_o(attention))   (attentions.py:308-309)
        o.row0 = l * ENC_ROWS_PER_LAYER + 3 * FD; o.ntiles = 4; o.chunk = 4; o.a_src = JA_TMA_TILE; o.att_row0 = 0; o.epi = JE_LN;
        o.bias = w.b_o; o.gamma = w.ln1_g; o.beta = w.ln1_b; o.res_in = p.res; o.res_out = p.res;
        JobDesc& a = add();   // fc1 + ReLU -> hidden scratch
        a.row0 = l * ENC_ROWS_PER_LAYER + 4 * FD; a.ntiles = 16; a.epi = JE_STORE; a.flags = JF_RELU | JF_WHOLE_TILES | JF_HIDDEN_DONE;
        a.bias = w.b_fc1; a.dst = p.hbuf; a.ld_dst = FDFF;
        JobDesc& b = add();   // LN(a + fc2(hidden)), padded rows zeroed (encoders.py:20), level output stored
        b.wmap = 1; b.row0 = l * FD; b.ntiles = 4; b.chunk = 4; b.kblocks = FDFF / BLOCK_K; b.a_src = JA_STREAM_HIDDEN; b.epi = JE_LN;
        b.bias = w.b_fc2; b.gamma = w.ln2_g; b.beta = w.ln2_b; b.res_in = p.res; b.res_out = p.res; b.zero_rows = d.row_mask;
        b.ln_out = static_cast<bf16*>(d.levels_out) + static_cast<size_t>(l) * d.level_stride;
        if (l + 1 < d.n_layers) qkv(l + 1);
        for (int i = 0; i < d.n_kv; ++i) {   // the decoder's cross K|V from this level's output, while it is resident
            if (d.kv_level[i] != l) continue;
            JobDesc& k = add();
            k.row0 = d.n_layers * ENC_ROWS_PER_LAYER + i * 2 * FD; k.ntiles = 8; k.epi = JE_STORE; k.bias = d.b_kv[i];
            k.dst = static_cast<bf16*>(d.kv_dst[i]); k.ld_dst = 2 * FD;
        }
    }
    CAP_REQUIRE(p.n_jobs < MAX_JOBS, "cap_enc_chain: job list overflow");
    return launch_chain(p, tiles, f->use_pairs, static_cast<cudaStream_t>(stream));
}
