// Beam-search state machine on the device (models/modules/beam_search.py:36-118 restated as three
// kernels, no host round trip inside the loop).
//
//   beam_rowpass_kernel   one CTA per beam row: EOS bookkeeping (seq_mask *= prev != eos), fused
//                         log-softmax (or pass-through log-probs), candidate value
//                         seq_logprob + word_logprob with the reference's -999 sentinel for
//                         finished beams, and the row-local top-`beam` under the stable-descending
//                         order (value desc, flat index asc).
//   beam_select_kernel    one CTA per image: merge the rows' candidates, pick the `beam` best,
//                         then reorder every per-beam state IN PLACE inside the image's group:
//                         seq_logprob, seq_mask, id / log-prob histories, next tokens and the
//                         ancestry table that replaces the reference's per-step KV-cache gather.
//   beam_finalize_kernel  final descending sort by seq_logprob + gather, int64 ids out.
#include "cap_common.cuh"

#include <atomic>

extern std::atomic<long long> g_cap_launches;

namespace {

constexpr int BEAM_MAX = 8;
constexpr int ROW_THREADS = 512;
constexpr float SENTINEL = -999.0f;  // models/modules/beam_search.py:54

struct Cand {
    float val;
    int idx;
};

__device__ __forceinline__ bool before(const Cand& a, const Cand& b) { return cand_before(a.val, a.idx, b.val, b.idx); }

__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.val = __shfl_xor_sync(0xffffffffu, c.val, o);
        other.idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
        if (before(other, c)) c = other;
    }
    return c;
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < ROW_THREADS / 32; ++w) r = fmaxf(r, red[w]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < ROW_THREADS / 32; ++w) r += red[w];
    __syncthreads();
    return r;
}

struct BeamDev {
    int32_t* tokens;      // [R]
    float* seq_logprob;   // [R]
    float* seq_mask;      // [R]
    int32_t* hist_ids;    // [R][T]
    float* hist_lp;       // [R][T]
    int32_t* ancestry;    // [T][R]
    int32_t* parents;     // [R]
    float* cand_val;      // [R][BEAM_MAX]
    float* cand_lp;       // [R][BEAM_MAX]
    int32_t* cand_idx;    // [R][BEAM_MAX]
    int batch, beam, max_len, vocab, eos;
};

__global__ void __launch_bounds__(ROW_THREADS)
beam_rowpass_kernel(const BeamDev st, const float* __restrict__ scores, int ld, int is_logprob, int t, int stage_smem) {
    pdl_prologue();
    extern __shared__ __align__(16) float row_smem[];
    __shared__ float red[ROW_THREADS / 32];
    __shared__ Cand warp_cand[ROW_THREADS / 32];
    __shared__ int s_winner;
    const int r = blockIdx.x;
    const int beam = st.beam, V = st.vocab;
    const int tid = threadIdx.x;

    float mask = st.seq_mask[r];
    if (t > 0) mask *= (st.tokens[r] != st.eos) ? 1.f : 0.f;
    const float seq_lp = st.seq_logprob[r];
    __syncthreads();  // everyone has read seq_mask before thread 0 rewrites it
    if (tid == 0) st.seq_mask[r] = mask;

    float* cval = st.cand_val + static_cast<size_t>(r) * BEAM_MAX;
    float* clp = st.cand_lp + static_cast<size_t>(r) * BEAM_MAX;
    int32_t* cidx = st.cand_idx + static_cast<size_t>(r) * BEAM_MAX;

    if (t == 0 && (r % beam) != 0) {  // cur_beam_size == 1: only beam 0 of each image competes
        if (tid < beam) { cval[tid] = -INFINITY; clp[tid] = 0.f; cidx[tid] = tid; }
        return;
    }
    if (mask == 0.f) {  // finished beam: column 0 keeps seq_logprob, every other column is -999
        if (tid < beam) {
            cval[tid] = (tid == 0) ? seq_lp : SENTINEL;
            clp[tid] = 0.f;  // word_logprob * seq_mask
            cidx[tid] = tid;
        }
        return;
    }

    const float* grow = scores + static_cast<size_t>(r) * ld;
    const float* row = grow;
    if (stage_smem) {
        for (int v = tid; v < V; v += ROW_THREADS) row_smem[v] = grow[v];
        __syncthreads();
        row = row_smem;
    }
    float mx = 0.f, log_sum = 0.f;
    if (!is_logprob) {
        float m = -INFINITY;
        for (int v = tid; v < V; v += ROW_THREADS) m = fmaxf(m, row[v]);
        mx = block_reduce_max(m, red);
        float s = 0.f;
        for (int v = tid; v < V; v += ROW_THREADS) s += __expf(row[v] - mx);
        log_sum = logf(block_reduce_sum(s, red));
    }

    // thread-local sorted top-`beam` over this thread's strided slice
    Cand best[BEAM_MAX];
#pragma unroll
    for (int i = 0; i < BEAM_MAX; ++i) { best[i].val = -INFINITY; best[i].idx = 0x7fffffff; }
    for (int v = tid; v < V; v += ROW_THREADS) {
        const float lp = is_logprob ? row[v] : (row[v] - mx) - log_sum;
        Cand c;
        c.val = seq_lp + lp;
        c.idx = v;
        if (before(c, best[BEAM_MAX - 1])) {  // static top-8 list keeps best[] in registers
#pragma unroll
            for (int i = BEAM_MAX - 1; i >= 0; --i) {
                if (i > 0 && before(c, best[i - 1])) {
                    best[i] = best[i - 1];
                } else {
                    best[i] = c;
                    break;
                }
            }
        }
    }
    // `beam` rounds of block-wide arg-best over the heads of the per-thread lists
    int head = 0;
    for (int round = 0; round < beam; ++round) {
        Cand mine;
        mine.val = -INFINITY;
        mine.idx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < BEAM_MAX; ++i)
            if (i == head && i < beam) mine = best[i];
        const Cand wb = warp_best(mine);
        if ((tid & 31) == 0) warp_cand[tid >> 5] = wb;
        __syncthreads();
        if (tid == 0) {
            Cand bb = warp_cand[0];
            for (int w = 1; w < ROW_THREADS / 32; ++w)
                if (before(warp_cand[w], bb)) bb = warp_cand[w];
            const float x = row[bb.idx];
            cval[round] = bb.val;
            cidx[round] = bb.idx;
            clp[round] = is_logprob ? x : (x - mx) - log_sum;
            s_winner = bb.idx;
        }
        __syncthreads();
        if (head < beam && mine.idx == s_winner && mine.val != -INFINITY) ++head;  // indices are unique per row
        __syncthreads();
    }
}

// Register-resident row pass (vocab <= 256*ITEMS): the row is read from HBM exactly once into
// statically indexed registers; max / sum-exp / `beam` rounds of block arg-best all run on registers
// with one __syncthreads per reduction (double-buffered partials).
template <int ITEMS>
__global__ void __launch_bounds__(ROW_THREADS, ITEMS <= 20 ? 2 : 1)
beam_rowpass_reg_kernel(const BeamDev st, const float* __restrict__ scores, int ld, int is_logprob, int t) {
    pdl_prologue();
    __shared__ float red[ROW_THREADS / 32];
    __shared__ float s_val[2][ROW_THREADS / 32];
    __shared__ int s_idx[2][ROW_THREADS / 32];
    __shared__ int s_win[BEAM_MAX];
    const int r = blockIdx.x;
    const int beam = st.beam, V = st.vocab;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float mask = st.seq_mask[r];
    if (t > 0) mask *= (st.tokens[r] != st.eos) ? 1.f : 0.f;
    const float seq_lp = st.seq_logprob[r];
    __syncthreads();
    if (tid == 0) st.seq_mask[r] = mask;

    float* cval = st.cand_val + static_cast<size_t>(r) * BEAM_MAX;
    float* clp = st.cand_lp + static_cast<size_t>(r) * BEAM_MAX;
    int32_t* cidx = st.cand_idx + static_cast<size_t>(r) * BEAM_MAX;
    if (t == 0 && (r % beam) != 0) {
        if (tid < beam) { cval[tid] = -INFINITY; clp[tid] = 0.f; cidx[tid] = tid; }
        return;
    }
    if (mask == 0.f) {
        if (tid < beam) { cval[tid] = (tid == 0) ? seq_lp : SENTINEL; clp[tid] = 0.f; cidx[tid] = tid; }
        return;
    }

    const float* row = scores + static_cast<size_t>(r) * ld;
    float x[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int v = tid + i * ROW_THREADS;
        x[i] = v < V ? __ldg(row + v) : -INFINITY;
    }
    float mx = 0.f, log_sum = 0.f;
    if (!is_logprob) {
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) m = fmaxf(m, x[i]);
        mx = block_reduce_max(m, red);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) s += __expf(x[i] - mx);  // padding lanes hold -inf -> 0
        log_sum = logf(block_reduce_sum(s, red));
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int v = tid + i * ROW_THREADS;
        const float lp = is_logprob ? x[i] : (x[i] - mx) - log_sum;
        x[i] = v < V ? seq_lp + lp : -INFINITY;  // candidate_logprob = seq_logprob + word_logprob
    }
    for (int round = 0; round < beam; ++round) {
        Cand mine;
        mine.val = -INFINITY;
        mine.idx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {  // indices ascend with i: strict > keeps the lowest index on ties
            if (x[i] > mine.val) { mine.val = x[i]; mine.idx = tid + i * ROW_THREADS; }
        }
        const Cand wb = warp_best(mine);
        const int buf = round & 1;
        if (lane == 0) { s_val[buf][warp] = wb.val; s_idx[buf][warp] = wb.idx; }
        __syncthreads();
        Cand bb;
        bb.val = s_val[buf][0];
        bb.idx = s_idx[buf][0];
#pragma unroll
        for (int w = 1; w < ROW_THREADS / 32; ++w) {
            Cand c;
            c.val = s_val[buf][w];
            c.idx = s_idx[buf][w];
            if (before(c, bb)) bb = c;
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (tid + i * ROW_THREADS == bb.idx) x[i] = -INFINITY;  // the owner retires the winner
        if (tid == 0) {
            const int widx = bb.idx == 0x7fffffff ? 0 : bb.idx;
            cval[round] = bb.val;
            cidx[round] = widx;
            s_win[round] = widx;
        }
    }
    __syncthreads();
    if (tid < beam) {  // the winners' word log-probs, recomputed from the row exactly as above
        const float xv = __ldg(row + s_win[tid]);
        clp[tid] = is_logprob ? xv : (xv - mx) - log_sum;
    }
}

// Row merge for the vocabulary GEMM's STATS epilogue (gemm_tcgen05.cu): one warp per beam row.
//   1. log-sum-exp of the row from the per-32-column-chunk (max, sum exp) pairs;
//   2. the `beam` chunks with the largest maxima, ties to the lower chunk -- the row's `beam` best logits
//      under (value desc, column asc) provably lie in them: a chunk ordered before another contributes
//      an element ordered before every element of the other;
//   3. only those chunks' logits are read back (lane = column) and the `beam` best candidates
//      seq_logprob + ((x - max) - log_sum) selected, with the row pass's EOS / -999 bookkeeping.
constexpr int MERGE_WARPS = 4;
constexpr int MERGE_CHUNKS_PER_LANE = 16;  // up to 512 chunks: vocab <= 16384

__global__ void __launch_bounds__(MERGE_WARPS * 32)
beam_chunkmerge_kernel(const BeamDev st, const float* __restrict__ logits, int ld, const float* __restrict__ part_ms,
                       int chunks, int t) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = st.batch * st.beam;
    const int r = blockIdx.x * MERGE_WARPS + warp;
    if (r >= R) return;
    const int beam = st.beam, V = st.vocab;

    float mask = st.seq_mask[r];
    if (t > 0) mask *= (st.tokens[r] != st.eos) ? 1.f : 0.f;
    const float seq_lp = st.seq_logprob[r];
    __syncwarp();
    if (lane == 0) st.seq_mask[r] = mask;
    float* cval = st.cand_val + static_cast<size_t>(r) * BEAM_MAX;
    float* clp = st.cand_lp + static_cast<size_t>(r) * BEAM_MAX;
    int32_t* cidx = st.cand_idx + static_cast<size_t>(r) * BEAM_MAX;
    if (t == 0 && (r % beam) != 0) {
        if (lane < beam) { cval[lane] = -INFINITY; clp[lane] = 0.f; cidx[lane] = lane; }
        return;
    }
    if (mask == 0.f) {
        if (lane < beam) { cval[lane] = (lane == 0) ? seq_lp : SENTINEL; clp[lane] = 0.f; cidx[lane] = lane; }
        return;
    }
    const float2* ms = reinterpret_cast<const float2*>(part_ms) + static_cast<size_t>(r) * chunks;
    float cm[MERGE_CHUNKS_PER_LANE];
    float mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int i = 0; i < MERGE_CHUNKS_PER_LANE; ++i) {
        const int chunk = lane + 32 * i;
        cm[i] = -INFINITY;
        if (chunk < chunks) cm[i] = ms[chunk].x;
        mx = fmaxf(mx, cm[i]);
    }
    mx = warp_max(mx);
#pragma unroll
    for (int i = 0; i < MERGE_CHUNKS_PER_LANE; ++i) {
        const int chunk = lane + 32 * i;
        if (chunk < chunks && cm[i] != -INFINITY) sum += ms[chunk].y * __expf(cm[i] - mx);
    }
    const float log_sum = logf(warp_sum(sum));

    // the `beam` best chunks, then this lane's column of each
    float cv[BEAM_MAX], lp[BEAM_MAX];
    int ci[BEAM_MAX];
#pragma unroll
    for (int k = 0; k < BEAM_MAX; ++k) {
        cv[k] = -INFINITY;
        lp[k] = 0.f;
        ci[k] = 0x7fffffff;
        if (k < beam) {
            Cand mine;
            mine.val = -INFINITY;
            mine.idx = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < MERGE_CHUNKS_PER_LANE; ++i)  // chunk ids ascend with i: strict > keeps the lower one
                if (cm[i] > mine.val) { mine.val = cm[i]; mine.idx = lane + 32 * i; }
            const Cand wb = warp_best(mine);
#pragma unroll
            for (int i = 0; i < MERGE_CHUNKS_PER_LANE; ++i)
                if (lane + 32 * i == wb.idx) cm[i] = -INFINITY;  // retired by its owner
            if (wb.idx != 0x7fffffff) {
                const int col = wb.idx * 32 + lane;
                if (col < V) {
                    const float x = __ldg(logits + static_cast<size_t>(r) * ld + col);
                    lp[k] = (x - mx) - log_sum;  // word_logprob, same formula as the row pass
                    cv[k] = seq_lp + lp[k];      // candidate_logprob
                    ci[k] = col;
                }
            }
        }
    }
    for (int round = 0; round < beam; ++round) {
        Cand mine;
        mine.val = -INFINITY;
        mine.idx = 0x7fffffff;
        float mine_lp = 0.f;
#pragma unroll
        for (int k = 0; k < BEAM_MAX; ++k) {
            if (cand_before(cv[k], ci[k], mine.val, mine.idx)) {
                mine.val = cv[k];
                mine.idx = ci[k];
                mine_lp = lp[k];
            }
        }
        const Cand wb = warp_best(mine);
        if (mine.idx == wb.idx && wb.idx != 0x7fffffff) {  // the owning lane publishes and retires it
            cval[round] = wb.val;
            cidx[round] = wb.idx;
            clp[round] = mine_lp;
#pragma unroll
            for (int k = 0; k < BEAM_MAX; ++k)
                if (ci[k] == wb.idx) { cv[k] = -INFINITY; ci[k] = 0x7fffffff; }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) beam_select_kernel(const BeamDev st, int t) {
    pdl_prologue();
    __shared__ Cand s_sel[BEAM_MAX];
    __shared__ float s_sel_lp[BEAM_MAX];
    __shared__ int s_sel_beam[BEAM_MAX];
    __shared__ float s_mask[BEAM_MAX];
    extern __shared__ __align__(16) uint8_t sel_smem[];
    const int b = blockIdx.x, beam = st.beam, V = st.vocab, T = st.max_len, R = st.batch * st.beam;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cur = (t == 0) ? 1 : beam;
    const int row0 = b * beam;

    if (warp == 0) {
        // up to BEAM_MAX*BEAM_MAX = 64 candidates: two per lane, flat index = src_beam*V + word
        Cand c[2];
        float lp[2];
        int src[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int slot = e * 32 + lane;
            const int kb = slot / beam, kc = slot % beam;
            c[e].val = -INFINITY; c[e].idx = 0x7fffffff; lp[e] = 0.f; src[e] = 0;
            if (kb < cur) {
                const size_t o = static_cast<size_t>(row0 + kb) * BEAM_MAX + kc;
                c[e].val = st.cand_val[o];
                c[e].idx = kb * V + st.cand_idx[o];
                lp[e] = st.cand_lp[o];
                src[e] = kb;
            }
        }
        for (int round = 0; round < beam; ++round) {
            Cand mine = before(c[1], c[0]) ? c[1] : c[0];
            const Cand wb = warp_best(mine);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (c[e].idx == wb.idx && c[e].val == wb.val && wb.idx != 0x7fffffff) {
                    s_sel[round] = wb;
                    s_sel_lp[round] = lp[e];
                    s_sel_beam[round] = src[e];
                    c[e].val = -INFINITY;
                    c[e].idx = 0x7fffffff;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- stage the image's old per-beam state, then rewrite it in the selected order ----
    int32_t* s_ids = reinterpret_cast<int32_t*>(sel_smem);                 // [beam][T]
    float* s_lps = reinterpret_cast<float*>(s_ids + beam * T);             // [beam][T]
    int32_t* s_anc = reinterpret_cast<int32_t*>(s_lps + beam * T);         // [T][beam]
    for (int i = tid; i < beam * t; i += blockDim.x) {
        const int k = i / t, tt = i % t;
        s_ids[k * T + tt] = st.hist_ids[static_cast<size_t>(row0 + k) * T + tt];
        s_lps[k * T + tt] = st.hist_lp[static_cast<size_t>(row0 + k) * T + tt];
        s_anc[tt * beam + k] = st.ancestry[static_cast<size_t>(tt) * R + row0 + k];
    }
    if (tid < beam) s_mask[tid] = st.seq_mask[row0 + tid];
    __syncthreads();
    for (int i = tid; i < beam * t; i += blockDim.x) {
        const int k = i / t, tt = i % t;
        const int src = s_sel_beam[k];
        st.hist_ids[static_cast<size_t>(row0 + k) * T + tt] = s_ids[src * T + tt];
        st.hist_lp[static_cast<size_t>(row0 + k) * T + tt] = s_lps[src * T + tt];
        st.ancestry[static_cast<size_t>(tt) * R + row0 + k] = s_anc[tt * beam + src];
    }
    if (tid < beam) {
        const int k = tid, src = s_sel_beam[k];
        const int word = s_sel[k].idx - src * V;
        const int row = row0 + k;
        st.hist_ids[static_cast<size_t>(row) * T + t] = word;
        st.hist_lp[static_cast<size_t>(row) * T + t] = s_sel_lp[k];
        st.ancestry[static_cast<size_t>(t) * R + row] = row0 + src;
        st.seq_logprob[row] = s_sel[k].val;
        st.seq_mask[row] = s_mask[src];
        st.tokens[row] = word;
        st.parents[row] = src;
    }
}

__global__ void beam_finalize_kernel(const BeamDev st, int out_size, int64_t* __restrict__ ids,
                                     float* __restrict__ logp) {
    pdl_prologue();
    __shared__ int order[BEAM_MAX];
    const int b = blockIdx.x, beam = st.beam, T = st.max_len;
    if (threadIdx.x == 0) {
        // stable descending selection sort of the image's beams by seq_logprob
        bool used[BEAM_MAX];
        for (int i = 0; i < BEAM_MAX; ++i) used[i] = false;
        for (int o = 0; o < beam; ++o) {
            int bi = -1;
            float bv = 0.f;
            for (int k = 0; k < beam; ++k) {
                if (used[k]) continue;
                const float v = st.seq_logprob[b * beam + k];
                if (bi < 0 || v > bv) { bi = k; bv = v; }
            }
            used[bi] = true;
            order[o] = bi;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < out_size * T; i += blockDim.x) {
        const int o = i / T, tt = i % T;
        const size_t src = static_cast<size_t>(b * beam + order[o]) * T + tt;
        const size_t dst = (static_cast<size_t>(b) * out_size + o) * T + tt;
        ids[dst] = st.hist_ids[src];
        logp[dst] = st.hist_lp[src];
    }
}

// out[r,:] = log_softmax(logits[r,:])  (decoders.py:123) -- standalone form for the module-level API
__global__ void __launch_bounds__(ROW_THREADS)
log_softmax_rows_kernel(const float* __restrict__ logits, int ld, float* __restrict__ out, int ldo, int V) {
    pdl_prologue();
    __shared__ float red[ROW_THREADS / 32];
    const float* row = logits + static_cast<size_t>(blockIdx.x) * ld;
    float* orow = out + static_cast<size_t>(blockIdx.x) * ldo;
    float m = -INFINITY;
    for (int v = threadIdx.x; v < V; v += ROW_THREADS) m = fmaxf(m, row[v]);
    const float mx = block_reduce_max(m, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += ROW_THREADS) s += __expf(row[v] - mx);
    const float log_sum = logf(block_reduce_sum(s, red));
    for (int v = threadIdx.x; v < V; v += ROW_THREADS) orow[v] = (row[v] - mx) - log_sum;
}

__global__ void beam_reset_kernel(const BeamDev st, int bos) {
    pdl_prologue();
    const int R = st.batch * st.beam;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) {
        st.tokens[i] = bos;
        st.seq_logprob[i] = 0.f;
        st.seq_mask[i] = 1.f;
        st.parents[i] = 0;
    }
    for (int j = i; j < R * st.max_len; j += gridDim.x * blockDim.x) {
        st.hist_ids[j] = 0;
        st.hist_lp[j] = 0.f;
        st.ancestry[j] = j % R;
    }
}

}  // namespace

struct cap_beam {
    BeamDev dev;
    int max_batch;
    void* arena;
};

extern "C" int cap_beam_create(int max_batch, int beam, int max_len, int vocab, int eos_idx, cap_beam** out) {
    CAP_REQUIRE(out != nullptr, "cap_beam_create: null out");
    CAP_REQUIRE(max_batch > 0 && beam > 0 && beam <= BEAM_MAX, "cap_beam_create: beam must be in [1,%d]", BEAM_MAX);
    CAP_REQUIRE(max_len > 0 && max_len <= 64 && vocab >= beam, "cap_beam_create: bad max_len/vocab");
    const size_t R = static_cast<size_t>(max_batch) * beam, T = max_len;
    const size_t words = R * 4 /*tokens,seq_lp,seq_mask,parents*/ + R * T * 3 + R * BEAM_MAX * 3;
    void* arena = nullptr;
    CAP_CHECK_CUDA(cudaMalloc(&arena, words * 4));
    cap_beam* h = new cap_beam();
    h->arena = arena;
    h->max_batch = max_batch;
    uint32_t* p = static_cast<uint32_t*>(arena);
    BeamDev& d = h->dev;
    d.tokens = reinterpret_cast<int32_t*>(p); p += R;
    d.seq_logprob = reinterpret_cast<float*>(p); p += R;
    d.seq_mask = reinterpret_cast<float*>(p); p += R;
    d.parents = reinterpret_cast<int32_t*>(p); p += R;
    d.hist_ids = reinterpret_cast<int32_t*>(p); p += R * T;
    d.hist_lp = reinterpret_cast<float*>(p); p += R * T;
    d.ancestry = reinterpret_cast<int32_t*>(p); p += R * T;
    d.cand_val = reinterpret_cast<float*>(p); p += R * BEAM_MAX;
    d.cand_lp = reinterpret_cast<float*>(p); p += R * BEAM_MAX;
    d.cand_idx = reinterpret_cast<int32_t*>(p); p += R * BEAM_MAX;
    d.batch = max_batch;
    d.beam = beam;
    d.max_len = max_len;
    d.vocab = vocab;
    d.eos = eos_idx;
    *out = h;
    return CAP_OK;
}

extern "C" int cap_beam_destroy(cap_beam* h) {
    if (!h) return CAP_OK;
    cudaFree(h->arena);
    delete h;
    return CAP_OK;
}

extern "C" int cap_beam_reset(cap_beam* h, int batch, int bos_idx, cap_stream_t stream) {
    CAP_REQUIRE(h != nullptr, "cap_beam_reset: null handle");
    CAP_REQUIRE(batch > 0 && batch <= h->max_batch, "cap_beam_reset: batch %d outside (0,%d]", batch, h->max_batch);
    h->dev.batch = batch;  // the [T][R] tables are laid out for the CURRENT R = batch*beam
    const int R = batch * h->dev.beam;
    // serialising launch: every later decode kernel may start early (PDL) and some prefetch the encode-time cross K|V
    // before their griddepcontrol.wait -- the encoder's GEMMs must have COMPLETED before the first decode kernel runs
    CAP_LAUNCH_SERIAL((beam_reset_kernel), (R + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), h->dev, bos_idx);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("beam_reset_kernel");
}

extern "C" int cap_beam_step(cap_beam* h, int t, const float* scores, int ld, int is_logprob, cap_stream_t stream) {
    CAP_REQUIRE(h && scores, "cap_beam_step: null pointer");
    CAP_REQUIRE(t >= 0 && t < h->dev.max_len, "cap_beam_step: step %d outside [0,%d)", t, h->dev.max_len);
    CAP_REQUIRE(ld >= h->dev.vocab, "cap_beam_step: ld < vocab");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const BeamDev& d = h->dev;
    const int R = d.batch * d.beam;
    const size_t row_bytes = static_cast<size_t>(d.vocab) * 4;
    const int stage = row_bytes <= 160 * 1024;
    static cap_device_once smem_once;
    CAP_PROPAGATE(cap_opt_in_smem(smem_once, beam_rowpass_kernel, 160 * 1024));
    const int items = (d.vocab + ROW_THREADS - 1) / ROW_THREADS;
    if (items <= 4)
        CAP_LAUNCH((beam_rowpass_reg_kernel<4>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else if (items <= 8)
        CAP_LAUNCH((beam_rowpass_reg_kernel<8>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else if (items <= 16)
        CAP_LAUNCH((beam_rowpass_reg_kernel<16>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else if (items <= 20)
        CAP_LAUNCH((beam_rowpass_reg_kernel<20>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else if (items <= 32)
        CAP_LAUNCH((beam_rowpass_reg_kernel<32>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else if (items <= 64)
        CAP_LAUNCH((beam_rowpass_reg_kernel<64>), R, ROW_THREADS, 0, s, d, scores, ld, is_logprob, t);
    else  // very large vocabularies: shared-memory / multi-pass variant
        CAP_LAUNCH((beam_rowpass_kernel), R, ROW_THREADS, stage ? row_bytes : 0, s, d, scores, ld, is_logprob, t, stage);
    CAP_PROPAGATE(cap_check_launch("beam_rowpass_kernel"));
    const size_t sel_smem = static_cast<size_t>(d.beam) * d.max_len * 12;
    CAP_LAUNCH((beam_select_kernel), d.batch, 128, sel_smem, s, d, t);
    g_cap_launches.fetch_add(2, std::memory_order_relaxed);
    return cap_check_launch("beam_select_kernel");
}

extern "C" int cap_beam_step_stats(cap_beam* h, int t, const float* logits, int ld, const float* part_ms, int chunks,
                                   cap_stream_t stream) {
    CAP_REQUIRE(h && logits && part_ms, "cap_beam_step_stats: null pointer");
    CAP_REQUIRE(t >= 0 && t < h->dev.max_len, "cap_beam_step_stats: step %d outside [0,%d)", t, h->dev.max_len);
    CAP_REQUIRE(chunks > 0 && chunks <= 32 * MERGE_CHUNKS_PER_LANE && chunks * 32 >= h->dev.vocab && ld >= h->dev.vocab,
                "cap_beam_step_stats: %d chunks unsupported for vocab %d", chunks, h->dev.vocab);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const BeamDev& d = h->dev;
    const int R = d.batch * d.beam;
    CAP_LAUNCH((beam_chunkmerge_kernel), (R + MERGE_WARPS - 1) / MERGE_WARPS, MERGE_WARPS * 32, 0, s, d, logits, ld, part_ms,
               chunks, t);
    CAP_PROPAGATE(cap_check_launch("beam_chunkmerge_kernel"));
    const size_t sel_smem = static_cast<size_t>(d.beam) * d.max_len * 12;
    CAP_LAUNCH((beam_select_kernel), d.batch, 128, sel_smem, s, d, t);
    g_cap_launches.fetch_add(2, std::memory_order_relaxed);
    return cap_check_launch("beam_select_kernel");
}

extern "C" int cap_beam_finalize(cap_beam* h, int out_size, int64_t* ids, float* logp, cap_stream_t stream) {
    CAP_REQUIRE(h && ids && logp, "cap_beam_finalize: null pointer");
    CAP_REQUIRE(out_size >= 1 && out_size <= h->dev.beam, "cap_beam_finalize: out_size %d outside [1,%d]", out_size,
                h->dev.beam);
    CAP_LAUNCH((beam_finalize_kernel), h->dev.batch, 128, 0, static_cast<cudaStream_t>(stream), h->dev, out_size, ids, logp);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("beam_finalize_kernel");
}

extern "C" int cap_log_softmax(const float* logits, int ld, float* out, int ldo, int rows, int V, cap_stream_t stream) {
    CAP_REQUIRE(logits && out && rows > 0 && V > 0 && ld >= V && ldo >= V, "cap_log_softmax: bad arguments");
    CAP_LAUNCH((log_softmax_rows_kernel), rows, ROW_THREADS, 0, static_cast<cudaStream_t>(stream), logits, ld, out, ldo, V);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("log_softmax_rows_kernel");
}

extern "C" const int32_t* cap_beam_tokens(cap_beam* h) { return h ? h->dev.tokens : nullptr; }
extern "C" const int32_t* cap_beam_ancestry(cap_beam* h) { return h ? h->dev.ancestry : nullptr; }
extern "C" const float* cap_beam_seq_logprob(cap_beam* h) { return h ? h->dev.seq_logprob : nullptr; }
extern "C" const int32_t* cap_beam_parents(cap_beam* h) { return h ? h->dev.parents : nullptr; }
