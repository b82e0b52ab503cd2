// Fused attention kernels (d_k = d_v = 64).
//
//  attention_kernel            softmax(q.k^T*scale + mask + log g | memory slots).v for one
//                              (batch, head, 64-query tile) per CTA; K/V staged once in shared
//                              memory (K rows padded to 33 words => conflict-free lane-per-key dots),
//                              fp32 softmax with warp shuffles.  Serves the encoder self-attention
//                              (all three variants), teacher-forced decoder attention and the
//                              decode-step cross-attention (the `beam` rows of an image are the
//                              queries of one CTA, so the image's K/V is read once, not per beam).
//  decode_self_attention_kernel  one warp per (row, head) over the beam-indirected KV cache.
//
// These are HBM/L2-bound (nq*nk*64 MACs per head is tiny); the tensor cores are reserved for the
// projections (gemm_tcgen05.cu).
#include "cap_common.cuh"
#include "tcgen05_ptx.cuh"

#include <algorithm>

#include <atomic>
#include <cstdlib>

extern std::atomic<long long> g_cap_launches;

namespace {

constexpr int HEAD_DIM = 64;
constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int Q_TILE = 64;
constexpr int MAX_KEYS = 160;
constexpr int KEY_CHUNKS = MAX_KEYS / 32;
constexpr int K_STRIDE = HEAD_DIM + 2;  // bf16 elements: 33 words per row

struct AttnDev {
    const bf16 *q, *k, *v;
    bf16* out;
    long long q_bs, k_bs, v_bs, o_bs;
    int ldq, ldk, ldv, ldo;
    const uint8_t* mask;
    long long mask_bs;
    int mask_qs;
    const float* geometry;
    const bf16 *mem_k, *mem_v;
    int n_mem, B, H, nq, nk;
    float scale;
    // adaptive attention (attentions.py:229-268): query i has one extra key = value = sentinel[b, i] ("language signal")
    const bf16* sentinel;
    long long s_bs;
    int lds;
};

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnDev a) {
    pdl_prologue();
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int nk_all = a.nk + a.n_mem;
    bf16* sk = reinterpret_cast<bf16*>(att_smem);                       // [nk_all][K_STRIDE]
    bf16* sv = sk + static_cast<size_t>(nk_all) * K_STRIDE;             // [nk_all][64]
    float* sq = reinterpret_cast<float*>(sv + static_cast<size_t>(nk_all) * HEAD_DIM);  // [warps][64]
    float* sp = sq + ATT_WARPS * HEAD_DIM;                              // [warps][MAX_KEYS]

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * Q_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- stage K (4-byte stores into the padded layout) and V (dense) ----
    const bf16* kbase = a.k + b * a.k_bs + h * HEAD_DIM;
    const bf16* vbase = a.v + b * a.v_bs + h * HEAD_DIM;
    for (int j = warp; j < nk_all; j += ATT_WARPS) {
        const bf16* krow = (j < a.nk) ? kbase + static_cast<size_t>(j) * a.ldk
                                      : a.mem_k + static_cast<size_t>(j - a.nk) * (a.H * HEAD_DIM) + h * HEAD_DIM;
        const bf16* vrow = (j < a.nk) ? vbase + static_cast<size_t>(j) * a.ldv
                                      : a.mem_v + static_cast<size_t>(j - a.nk) * (a.H * HEAD_DIM) + h * HEAD_DIM;
        reinterpret_cast<bf162*>(sk + j * K_STRIDE)[lane] = reinterpret_cast<const bf162*>(krow)[lane];
        reinterpret_cast<bf162*>(sv + j * HEAD_DIM)[lane] = reinterpret_cast<const bf162*>(vrow)[lane];
    }
    __syncthreads();

    float* myq = sq + warp * HEAD_DIM;
    float* myp = sp + warp * MAX_KEYS;
    const int q_end = min(q0 + Q_TILE, a.nq);
    for (int i = q0 + warp; i < q_end; i += ATT_WARPS) {
        const bf16* qrow = a.q + b * a.q_bs + static_cast<size_t>(i) * a.ldq + h * HEAD_DIM;
        const float2 qv = __bfloat1622float2(reinterpret_cast<const bf162*>(qrow)[lane]);
        myq[2 * lane] = qv.x * a.scale;
        myq[2 * lane + 1] = qv.y * a.scale;
        __syncwarp();

        const uint8_t* mrow = a.mask ? a.mask + b * a.mask_bs + static_cast<size_t>(i) * a.mask_qs : nullptr;
        const float* grow = a.geometry
                                ? a.geometry + ((static_cast<size_t>(b) * a.H + h) * a.nq + i) * a.nk
                                : nullptr;
        float s[KEY_CHUNKS];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < KEY_CHUNKS; ++c) {
            const int j = c * 32 + lane;
            s[c] = -INFINITY;
            if (j < nk_all) {
                const bf162* kr = reinterpret_cast<const bf162*>(sk + j * K_STRIDE);
                float acc = 0.f;
#pragma unroll
                for (int d2 = 0; d2 < HEAD_DIM / 2; ++d2) {
                    const float2 kk = __bfloat1622float2(kr[d2]);
                    acc = fmaf(myq[2 * d2], kk.x, acc);
                    acc = fmaf(myq[2 * d2 + 1], kk.y, acc);
                }
                if (j < a.nk) {
                    if (mrow && mrow[j]) acc = -INFINITY;
                    if (grow) acc += logf(fmaxf(grow[j], 1e-6f));
                }
                s[c] = acc;
            }
            mx = fmaxf(mx, s[c]);
        }
        mx = warp_max(mx);
        // adaptive attention: the query's own sentinel is one more (never masked) column of the softmax
        float2 sent = make_float2(0.f, 0.f);
        float s_logit = -INFINITY;
        if (a.sentinel != nullptr) {
            const bf16* srow = a.sentinel + b * a.s_bs + static_cast<size_t>(i) * a.lds + h * HEAD_DIM;
            sent = __bfloat1622float2(reinterpret_cast<const bf162*>(srow)[lane]);
            s_logit = warp_sum(myq[2 * lane] * sent.x + myq[2 * lane + 1] * sent.y);
            mx = fmaxf(mx, s_logit);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < KEY_CHUNKS; ++c) {
            const int j = c * 32 + lane;
            const float p = (j < nk_all && mx != -INFINITY) ? __expf(s[c] - mx) : 0.f;
            if (j < nk_all) myp[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float p_sent = a.sentinel != nullptr ? __expf(s_logit - mx) : 0.f;
        sum += p_sent;
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        __syncwarp();
        float o0 = p_sent * sent.x, o1 = p_sent * sent.y;
        for (int j = 0; j < nk_all; ++j) {
            const float p = myp[j];
            const float2 vv = __bfloat1622float2(reinterpret_cast<const bf162*>(sv + j * HEAD_DIM)[lane]);
            o0 = fmaf(p, vv.x, o0);
            o1 = fmaf(p, vv.y, o1);
        }
        bf16* orow = a.out + b * a.o_bs + static_cast<size_t>(i) * a.ldo + h * HEAD_DIM;
        reinterpret_cast<bf162*>(orow)[lane] = __floats2bfloat162_rn(o0 * inv, o1 * inv);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------- tensor-core variant
// Same contract as attention_kernel for nk + n_mem <= 128.  One CTA = 4 warps = 64 query rows of one
// (batch, head); each warp owns a 16-row tile and keeps S = q.k^T, the softmax and P entirely in
// registers (FlashAttention-2 fragment reuse: the fp32 accumulator layout of two 8-key tiles is the bf16
// A-operand layout of one 16-key step), with K, V^T and Q staged once in shared memory.  The contraction
// runs on mma.sync m16n8k16 (bf16 in, fp32 accumulate): at 49..128 keys x 64 dims per head the tiles are
// far below tcgen05's 128-row atoms, so the warp-level tensor path is the right tool here; the
// projections around it are the tcgen05 GEMMs.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

constexpr int MMA_Q_TILE = 64;
constexpr int MMA_PITCH = HEAD_DIM + 8;  // bf16 elements: 36 words per row => conflict-free fragment loads

template <int NT>  // number of 8-key tiles: 8 (<= 64 keys) or 16 (<= 128 keys)
__global__ void __launch_bounds__(128) attention_mma_kernel(const AttnDev a) {
    if (threadIdx.x == 0) flight_mark(FK_ATTENTION, 0);
    pdl_prologue();
    constexpr int NKP = NT * 8;
    constexpr int VP = NKP + 8;  // V^T pitch (bf16): (NKP+8)/2 words == 4 mod 32 => conflict-free too
    __shared__ __align__(16) bf16 sK[NKP * MMA_PITCH];
    __shared__ __align__(16) bf16 sVt[HEAD_DIM * VP];
    __shared__ __align__(16) bf16 sQ[MMA_Q_TILE * MMA_PITCH];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MMA_Q_TILE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk_all = a.nk + a.n_mem;
    const int hd = a.H * HEAD_DIM;

    const bf16* kbase = a.k + b * a.k_bs + h * HEAD_DIM;
    const bf16* vbase = a.v + b * a.v_bs + h * HEAD_DIM;
    const bf162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    for (int idx = tid; idx < NKP * 32; idx += 128) {
        const int j = idx >> 5, dp = idx & 31;
        bf162 kk = zero2, vv = zero2;
        if (j < nk_all) {
            const bf16* krow = (j < a.nk) ? kbase + static_cast<size_t>(j) * a.ldk
                                          : a.mem_k + static_cast<size_t>(j - a.nk) * hd + h * HEAD_DIM;
            const bf16* vrow = (j < a.nk) ? vbase + static_cast<size_t>(j) * a.ldv
                                          : a.mem_v + static_cast<size_t>(j - a.nk) * hd + h * HEAD_DIM;
            kk = reinterpret_cast<const bf162*>(krow)[dp];
            vv = reinterpret_cast<const bf162*>(vrow)[dp];
        }
        *reinterpret_cast<bf162*>(sK + j * MMA_PITCH + 2 * dp) = kk;
        sVt[(2 * dp) * VP + j] = vv.x;
        sVt[(2 * dp + 1) * VP + j] = vv.y;
    }
    for (int idx = tid; idx < MMA_Q_TILE * 32; idx += 128) {
        const int i = idx >> 5, dp = idx & 31;
        bf162 qq = zero2;
        if (q0 + i < a.nq)
            qq = reinterpret_cast<const bf162*>(a.q + b * a.q_bs + static_cast<size_t>(q0 + i) * a.ldq + h * HEAD_DIM)[dp];
        *reinterpret_cast<bf162*>(sQ + i * MMA_PITCH + 2 * dp) = qq;
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16;
    if (q0 + r0 >= a.nq) return;  // whole 16-row tile out of range (warp-uniform)
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const bf16* base = sQ + (r0 + g) * MMA_PITCH + ks * 16 + 2 * t;
        qa[ks][0] = *reinterpret_cast<const uint32_t*>(base);
        qa[ks][1] = *reinterpret_cast<const uint32_t*>(base + 8 * MMA_PITCH);
        qa[ks][2] = *reinterpret_cast<const uint32_t*>(base + 8);
        qa[ks][3] = *reinterpret_cast<const uint32_t*>(base + 8 * MMA_PITCH + 8);
    }
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const bf16* base = sK + (nt * 8 + g) * MMA_PITCH + ks * 16 + 2 * t;
            mma_bf16_16816(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(base),
                           *reinterpret_cast<const uint32_t*>(base + 8));
        }
    }
    // scale, mask, geometry bias; rows ra = g, rb = g + 8 of this warp's tile
    const int ra = q0 + r0 + g, rb = ra + 8;
    const bool va = ra < a.nq, vb = rb < a.nq;
    const uint8_t* mra = (a.mask && va) ? a.mask + b * a.mask_bs + static_cast<size_t>(ra) * a.mask_qs : nullptr;
    const uint8_t* mrb = (a.mask && vb) ? a.mask + b * a.mask_bs + static_cast<size_t>(rb) * a.mask_qs : nullptr;
    const float* gra = (a.geometry && va) ? a.geometry + ((static_cast<size_t>(b) * a.H + h) * a.nq + ra) * a.nk : nullptr;
    const float* grb = (a.geometry && vb) ? a.geometry + ((static_cast<size_t>(b) * a.H + h) * a.nq + rb) * a.nk : nullptr;
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * t + (e & 1);
            const bool second = e >= 2;
            float x = s[nt][e] * a.scale;
            if (col >= nk_all) {
                x = -INFINITY;
            } else if (col < a.nk) {
                const uint8_t* mr = second ? mrb : mra;
                const float* gr = second ? grb : gra;
                if (mr && mr[col]) x = -INFINITY;
                if (gr) x += logf(fmaxf(gr[col], 1e-6f));
            }
            s[nt][e] = x;
            if (second) mb = fmaxf(mb, x); else ma = fmaxf(ma, x);
        }
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float suma = 0.f, sumb = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool second = e >= 2;
            const float m = second ? mb : ma;
            const float p = (m == -INFINITY) ? 0.f : __expf(s[nt][e] - m);
            s[nt][e] = p;
            if (second) sumb += p; else suma += p;
        }
    }
    suma += __shfl_xor_sync(0xffffffffu, suma, 1);
    suma += __shfl_xor_sync(0xffffffffu, suma, 2);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 1);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);
    // O = P.V : two 8-key accumulator tiles form one 16-key A operand
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const bf16* base = sVt + (dt * 8 + g) * VP + kk * 16 + 2 * t;
            mma_bf16_16816(o[dt], pa, *reinterpret_cast<const uint32_t*>(base),
                           *reinterpret_cast<const uint32_t*>(base + 8));
        }
    }
    const float inva = suma > 0.f ? 1.f / suma : 0.f;
    const float invb = sumb > 0.f ? 1.f / sumb : 0.f;
    bf16* oa = a.out + b * a.o_bs + static_cast<size_t>(ra) * a.ldo + h * HEAD_DIM;
    bf16* ob = a.out + b * a.o_bs + static_cast<size_t>(rb) * a.ldo + h * HEAD_DIM;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + 2 * t;
        if (va) *reinterpret_cast<bf162*>(oa + col) = __floats2bfloat162_rn(o[dt][0] * inva, o[dt][1] * inva);
        if (vb) *reinterpret_cast<bf162*>(ob + col) = __floats2bfloat162_rn(o[dt][2] * invb, o[dt][3] * invb);
    }
    if (threadIdx.x == 0) flight_mark(FK_ATTENTION, 1);
}

// ------------------------------------------------------------------------------ decode self-attn
constexpr int DEC_WARPS = 4;
constexpr int DEC_MAX_T = 64;

__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_self_attention_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ ancestry,
                             const uint8_t* __restrict__ padflag, bf16* __restrict__ out, int ldo, int t, int R,
                             int H, float scale) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * DEC_WARPS + warp;
    if (item >= R * H) return;
    const int r = item / H, h = item % H;
    const int hd = H * HEAD_DIM;
    const size_t step_stride = static_cast<size_t>(R) * 3 * hd;
    const int nkeys = t + 1;

    // whole query vector in registers (broadcast 16-byte loads), packed bf16x2
    bf162 qreg[HEAD_DIM / 2];
    {
        const bf16x8* qp = reinterpret_cast<const bf16x8*>(qkv + t * step_stride + static_cast<size_t>(r) * 3 * hd +
                                                          h * HEAD_DIM);
#pragma unroll
        for (int c = 0; c < HEAD_DIM / 8; ++c) {
            const bf16x8 v = qp[c];
#pragma unroll
            for (int i = 0; i < 4; ++i) qreg[c * 4 + i] = v.v[i];
        }
    }
    float s[2];
    int slot[2];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int j = c * 32 + lane;
        s[c] = -INFINITY;
        slot[c] = 0;
        if (j < nkeys) {
            const int sl = (j == t) ? r : ancestry[static_cast<size_t>(j) * R + r];
            slot[c] = sl;
            const bf16x8* kp = reinterpret_cast<const bf16x8*>(qkv + j * step_stride + static_cast<size_t>(sl) * 3 * hd +
                                                              hd + h * HEAD_DIM);
            float acc = 0.f;
#pragma unroll
            for (int cc = 0; cc < HEAD_DIM / 8; ++cc) {
                const bf16x8 kv = kp[cc];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 kk = __bfloat1622float2(kv.v[i]);
                    const float2 qq = __bfloat1622float2(qreg[cc * 4 + i]);
                    acc = fmaf(qq.x, kk.x, acc);
                    acc = fmaf(qq.y, kk.y, acc);
                }
            }
            acc *= scale;
            if (padflag[static_cast<size_t>(j) * R + sl]) acc = -INFINITY;
            s[c] = acc;
        }
        mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float p[2], sum = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int j = c * 32 + lane;
        p[c] = (j < nkeys && mx != -INFINITY) ? __expf(s[c] - mx) : 0.f;
        sum += p[c];
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nkeys; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p[j >> 5], j & 31);
        const int sl = __shfl_sync(0xffffffffu, slot[j >> 5], j & 31);
        const bf162* vp = reinterpret_cast<const bf162*>(qkv + j * step_stride + static_cast<size_t>(sl) * 3 * hd +
                                                        2 * hd + h * HEAD_DIM);
        const float2 vv = __bfloat1622float2(vp[lane]);
        o0 = fmaf(pj, vv.x, o0);
        o1 = fmaf(pj, vv.y, o1);
    }
    bf16* orow = out + static_cast<size_t>(r) * ldo + h * HEAD_DIM;
    reinterpret_cast<bf162*>(orow)[lane] = __floats2bfloat162_rn(o0 * inv, o1 * inv);
}

// Wide variant for H*64 == 32*EPL (H = 4, 8, 16): ONE warp per row covers all heads.  Lane l owns
// EPL consecutive elements of the H*64-wide q/k/v rows (head = l*EPL/64), so every key costs one fully
// coalesced row read for K and one for V with all 32 lanes busy at any step t; per-head scores are
// reduced across the 64/EPL lanes of the head with shuffles; softmax is online (single pass over keys).
// Warps per CTA come from the launch (4 by default).  Fat CTAs (16 warps) put the same rows on a quarter of the SMs:
// the fused decode chains (decode_fused.cu) need SMs that are otherwise EMPTY, and a thin attention kernel spread
// over all 148 SMs keeps every one of them "dirty" for its whole duration.
constexpr int DEC_WARPS_MAX = 16;
template <int EPL>
__global__ void __launch_bounds__(DEC_WARPS_MAX * 32)
decode_self_attention_wide_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ ancestry,
                                  const uint8_t* __restrict__ padflag, bf16* __restrict__ out, int ldo, int t, int R,
                                  float scale) {
    if (threadIdx.x == 0) flight_mark(FK_SELF_ATTENTION, 0);
    pdl_prologue();
    constexpr int LANES_PER_HEAD = HEAD_DIM / EPL;
    constexpr int VEC = EPL / 8;  // 16-byte vectors per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= R) return;
    const int hd = 32 * EPL;
    const size_t row_stride = static_cast<size_t>(3) * hd;
    const size_t step_stride = static_cast<size_t>(R) * row_stride;
    const int nkeys = t + 1;

    float q[EPL];
    {
        const bf16x8* qp = reinterpret_cast<const bf16x8*>(qkv + t * step_stride + r * row_stride + lane * EPL);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            unpack8(qp[c], q + c * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[c * 8 + i] *= scale;
        }
    }
    // this lane's key slot / pad flag for key index == lane (and lane + 32), fetched once
    int my_slot[2];
    bool my_pad[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int j = c * 32 + lane;
        my_slot[c] = 0;
        my_pad[c] = true;
        if (j < nkeys) {
            my_slot[c] = (j == t) ? r : ancestry[static_cast<size_t>(j) * R + r];
            my_pad[c] = padflag[static_cast<size_t>(j) * R + my_slot[c]] != 0;
        }
    }
    float m = -INFINITY, l = 0.f;
    float acc[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = 0.f;

    constexpr int UNROLL = 4;
    for (int j0 = 0; j0 < nkeys; j0 += UNROLL) {
        bf16x8 kreg[UNROLL][VEC], vreg[UNROLL][VEC];
        bool pad[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int j = min(j0 + u, nkeys - 1);  // tail iterations re-read the last key, masked below
            const int sl = __shfl_sync(0xffffffffu, my_slot[j >> 5], j & 31);
            pad[u] = __shfl_sync(0xffffffffu, static_cast<int>(my_pad[j >> 5]), j & 31) != 0 || (j0 + u >= nkeys);
            const bf16* base = qkv + j * step_stride + sl * row_stride + lane * EPL;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                // cache rows are read once per step: streaming loads (evict-first in L2)
                const uint4 ku = __ldcs(reinterpret_cast<const uint4*>(base + hd) + c);
                const uint4 vu = __ldcs(reinterpret_cast<const uint4*>(base + 2 * hd) + c);
                kreg[u][c] = *reinterpret_cast<const bf16x8*>(&ku);
                vreg[u][c] = *reinterpret_cast<const bf16x8*>(&vu);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                float kf[8];
                unpack8(kreg[u][c], kf);
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(q[c * 8 + i], kf[i], s);
            }
#pragma unroll
            for (int o = LANES_PER_HEAD / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (pad[u]) continue;  // masked key (uniform across the warp)
            const float m_new = fmaxf(m, s);
            const float corr = __expf(m - m_new);  // exp(-inf) = 0 on the first live key
            const float p = __expf(s - m_new);
            l = l * corr + p;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                float vf[8];
                unpack8(vreg[u][c], vf);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c * 8 + i] = acc[c * 8 + i] * corr + p * vf[i];
            }
            m = m_new;
        }
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = out + static_cast<size_t>(r) * ldo + lane * EPL;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[c * 8 + i] * inv;
        reinterpret_cast<bf16x8*>(orow)[c] = pack8(o);
    }
    if (threadIdx.x == 0) flight_mark(FK_SELF_ATTENTION, 1);
}

// ---------------------------------------------------------------------------- decode cross-attention
// Decode-step cross-attention on the warp-level tensor path (H = 8): one warp per head; an image's K|V rows are
// bulk-copied (one cp.async.bulk per 2 KB row) into shared-memory rows of pitch 2 KB + 16 B, which makes both fragment
// access patterns conflict-free: K as the col-major B operand of S = Q.K^T by 32-bit loads (lanes of a quad-group walk
// rows, pitch = 4 banks), V as the B operand of O = P.V by ldmatrix.trans (8 rows x 16 B, pitch = one 16-byte bank
// group).  The image's beams are rows 0..beams-1 of the m16 tile (the other rows are zero: most of the MMA is padding,
// still ~7x fewer instructions than CUDA-core arithmetic, whose per-key unpack / FMA / exp chain is issue-bound -- the
// round-1 kernels of that kind measured 23.3 us per launch at the bench shape against 16.4 us); keys past n read one
// shared zero row.  S accumulators become the P operand in registers (bf16), softmax in fp32 in the log2 domain.
constexpr int XT_PITCH = 2048 + 16;   // bytes per staged K|V row

__device__ __forceinline__ uint32_t xs_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// Streamed: a first version ran one CTA per image -- load 100 KB, wait, compute, exit: nothing overlapped inside a CTA,
// and its 256 CTAs held 200 KB of shared memory on 128 SMs for the whole launch.  Here a CTA is a small pipeline: a PRODUCER warp streams image after image into a ring of shared-memory stages (one cp.async.bulk
// per 2 KB K|V row, one mbarrier phase per image), eight CONSUMER warps (one per head) run the same mma.sync
// arithmetic on the stage that has landed while the next image is in flight, and hand the stage back through an
// "empty" barrier.  A quarter of the CTAs (4 images each) then move the same bytes: the launch costs a fraction of
// the SM-time, which is what the pipelined regime (32 batches in flight) is short of.
constexpr int XA_CONSUMERS = 8;                       // one warp per head
constexpr int XA_THREADS = (XA_CONSUMERS + 1) * 32;   // + the producer warp
constexpr int XA_MAX_STAGES = 4;
constexpr int XA_IMAGES_PER_CTA = 4;

template <int NT>   // 8-key tiles: 7 (n <= 56) or 13 (n <= 104)
__global__ void __launch_bounds__(XA_THREADS, 1)
decode_cross_attention_stream_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ kv,
                                     const uint8_t* __restrict__ key_mask, bf16* __restrict__ out, int ldo, int beams, int n,
                                     float scale, int n_images, int stages, uint32_t stage_bytes, int levels,
                                     size_t kv_level_stride, size_t out_level_stride) {
    // `levels` > 1 (meshed decoder, decoders.py:55-57): the same queries attend to the K|V of several encoder levels;
    // virtual image v = level * n_images + image reads kv + level * kv_level_stride, writes out + level * out_level_stride
    pdl_launch_dependents();
    const int n_virtual = n_images * levels;
    extern __shared__ __align__(128) uint8_t xa_smem[];   // [stages][n rows at XT_PITCH], zero row, [stages][NT*8] mask, barriers
    uint8_t* zero_row = xa_smem + static_cast<size_t>(stages) * stage_bytes;
    uint8_t* smask = zero_row + XT_PITCH;
    uint64_t* full = reinterpret_cast<uint64_t*>(smask + XA_MAX_STAGES * NT * 8);
    uint64_t* empty = full + XA_MAX_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            cap_ptx::mbar_init(&full[s], 1);
            cap_ptx::mbar_init(&empty[s], XA_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < XT_PITCH / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(zero_row)[i] = 0u;
    __syncthreads();

    if (warp == XA_CONSUMERS) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) flight_mark(FK_CROSS_PRODUCER, 0);
        // K|V were projected at encode time and a serialising launch separates encode from decode: safe to stream
        // before (without) the PDL wait.  Read once per step: evict-first, the weights keep their place in L2.
        const uint64_t stream_policy = cap_ptx::l2_policy_evict_first();
        int it = 0;
        for (int vi = blockIdx.x; vi < n_virtual; vi += gridDim.x, ++it) {
            const int img = vi % n_images, lvl = vi / n_images;
            const int s = it % stages;
            const uint32_t use = static_cast<uint32_t>(it / stages);
            if (use > 0) cap_ptx::mbar_wait(&empty[s], (use - 1) & 1);   // every consumer is done with the stage's last image
            uint8_t* mk = smask + s * NT * 8;
            for (int j = lane; j < NT * 8; j += 32)
                mk[j] = (j >= n || (key_mask != nullptr && key_mask[static_cast<size_t>(img) * n + j])) ? 1 : 0;
            __syncwarp();
            if (lane == 0) cap_ptx::mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(n) * 2048u);   // release: mask visible
            __syncwarp();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(kv + lvl * kv_level_stride) + static_cast<size_t>(img) * n * 2048;
            uint8_t* dst = xa_smem + static_cast<size_t>(s) * stage_bytes;
            for (int r = lane; r < n; r += 32)
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                        xs_smem_u32(dst + static_cast<size_t>(r) * XT_PITCH)),
                    "l"(reinterpret_cast<uint64_t>(src + static_cast<size_t>(r) * 2048)), "r"(2048u), "r"(xs_smem_u32(&full[s])),
                    "l"(stream_policy)
                    : "memory");
        }
        if (lane == 0) flight_mark(FK_CROSS_PRODUCER, 1);
        return;
    }

    // ---------------------------------------------------------------------- consumers: warp = head
    const int g = lane >> 2, t = lane & 3;
    const int h = warp;
    const float sc = scale * 1.4426950408889634f;
    const int mi = lane >> 3, mr = lane & 7;   // ldmatrix: this lane addresses row mr of 8x8 matrix mi of an x4 load
    if (threadIdx.x == 0) flight_mark(FK_CROSS_CONSUMER, 0);
    pdl_wait();   // q comes from the previous kernel; out is read by nothing that is still running
    if (threadIdx.x == 0) flight_mark(FK_CROSS_CONSUMER_READY, 0);
    // A operand of S = Q.K^T: rows g < beams of the image's queries, head h; rows 8..15 of the m16 tile are zero
    auto load_q = [&](int img, uint32_t (&qa)[4][4]) {
        const bf16* qrow = q + static_cast<size_t>(img * beams + (g < beams ? g : 0)) * ldq + h * HEAD_DIM + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t lo = *reinterpret_cast<const uint32_t*>(qrow + ks * 16);
            const uint32_t hi = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8);
            qa[ks][0] = g < beams ? lo : 0u;
            qa[ks][1] = 0u;
            qa[ks][2] = g < beams ? hi : 0u;
            qa[ks][3] = 0u;
        }
    };
    uint32_t qa[4][4];
    if (blockIdx.x < n_virtual) load_q(blockIdx.x % n_images, qa);
    int it = 0;
    for (int vi = blockIdx.x; vi < n_virtual; vi += gridDim.x, ++it) {
        const int img = vi % n_images, lvl = vi / n_images;
        const int s = it % stages;
        const uint32_t use = static_cast<uint32_t>(it / stages);
        uint32_t qn[4][4];   // the next image's queries are requested before this image's arithmetic
        const int next = vi + gridDim.x;
        if (next < n_virtual) load_q(next % n_images, qn);
        cap_ptx::mbar_wait(&full[s], use & 1);
        const uint8_t* st = xa_smem + static_cast<size_t>(s) * stage_bytes;
        const uint8_t* mk = smask + s * NT * 8;
        // S = Q.K^T
        float sv[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            sv[nt][0] = sv[nt][1] = sv[nt][2] = sv[nt][3] = 0.f;
            const int key = nt * 8 + g;
            const uint8_t* krow = (key < n ? st + static_cast<size_t>(key) * XT_PITCH : zero_row) + h * (HEAD_DIM * 2) + 4 * t;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                mma_bf16_16816(sv[nt], qa[ks], *reinterpret_cast<const uint32_t*>(krow + ks * 32),
                               *reinterpret_cast<const uint32_t*>(krow + ks * 32 + 16));
        }
        // softmax over the keys of row g (log2 domain); a quad holds one row
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float x = mk[nt * 8 + 2 * t + e] ? -INFINITY : sv[nt][e] * sc;
                sv[nt][e] = x;
                mx = fmaxf(mx, x);
            }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float p = (mx == -INFINITY) ? 0.f : fast_exp2(sv[nt][e] - mx);
                sv[nt][e] = p;
                sum += p;
            }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        // O = P.V : two 8-key score tiles are the A operand of one 16-key step; V fragments by ldmatrix.trans
        float o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < (NT + 1) / 2; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(sv[2 * kk][0], sv[2 * kk][1]);
            pa[1] = 0u;
            pa[2] = (2 * kk + 1 < NT) ? pack_bf16x2(sv[2 * kk + 1 < NT ? 2 * kk + 1 : 0][0], sv[2 * kk + 1 < NT ? 2 * kk + 1 : 0][1]) : 0u;
            pa[3] = 0u;
            const int key = kk * 16 + (mi & 1) * 8 + mr;
            const uint8_t* vrow = (key < n ? st + static_cast<size_t>(key) * XT_PITCH : zero_row) + 1024 + h * (HEAD_DIM * 2) +
                                  (mi >> 1) * 16;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, xs_smem_u32(vrow + dp * 32));
                mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
                mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
            }
        }
        __syncwarp();   // every lane's shared-memory reads of the stage are done
        if (lane == 0) cap_ptx::mbar_arrive(&empty[s]);
        if (g < beams) {
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
            bf16* orow = out + lvl * out_level_stride + static_cast<size_t>(img * beams + g) * ldo + h * HEAD_DIM + 2 * t;
#pragma unroll
            for (int dt = 0; dt < 8; ++dt)
                *reinterpret_cast<bf162*>(orow + dt * 8) = __floats2bfloat162_rn(o[dt][0] * inv, o[dt][1] * inv);
        }
        if (next < n_virtual) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qa[ks][0] = qn[ks][0];
                qa[ks][2] = qn[ks][2];
            }
        }
    }
    if (threadIdx.x == 0) {
        flight_mark(FK_CROSS_CONSUMER, 1);
        flight_mark(FK_CROSS_CONSUMER_READY, 1);
    }
}

template <int NT>
int launch_cross_stream(const bf16* q, int ldq, const bf16* kv, const uint8_t* key_mask, bf16* out, int ldo, int B, int beams,
                        int n, float scale, cudaStream_t stream, int levels = 1, size_t kv_level_stride = 0,
                        size_t out_level_stride = 0) {
    const uint32_t stage_bytes = static_cast<uint32_t>(n) * XT_PITCH;
    const size_t fixed = XT_PITCH + XA_MAX_STAGES * NT * 8 + 2 * XA_MAX_STAGES * 8;
    const size_t budget = 226 * 1024;
    int stages = static_cast<int>((budget - fixed) / stage_bytes);
    if (stages < 1) return cap_set_error(CAP_ERR_INVALID, "cap_decode_cross_attention: %d keys do not fit shared memory", n);
    stages = std::min(stages, XA_MAX_STAGES);
    const int per_cta = stages > 1 ? XA_IMAGES_PER_CTA : 1;
    const int n_virtual = B * levels;
    const int grid = std::max(1, (n_virtual + per_cta - 1) / per_cta);
    stages = std::min(stages, (n_virtual + grid - 1) / grid);   // never more stages than images per CTA
    const size_t smem = static_cast<size_t>(stages) * stage_bytes + fixed;
    static cap_device_once smem_once;
    CAP_PROPAGATE(cap_opt_in_smem(smem_once, decode_cross_attention_stream_kernel<NT>, 227 * 1024));
    CAP_PROPAGATE(cap_ptx::install_fault_buffer());
    CAP_LAUNCH((decode_cross_attention_stream_kernel<NT>), grid, XA_THREADS, smem, stream, q, ldq, kv, key_mask, out, ldo, beams,
               n, scale, B, stages, stage_bytes, levels, kv_level_stride, out_level_stride);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("decode_cross_attention_stream_kernel");
}

// Encoder self-attention with the whole image in shared memory (plain scaled dot-product,
// H = 8, fused q|k|v rows, n <= 64): one CTA per image, one warp per head.  The image's n rows of q|k|v (3 KB each,
// contiguous in the fused projection output) are bulk-copied once into rows of pitch 3 KB + 16 B; every warp then
// walks the 16-query tiles of its head: S = Q.K^T with A and B fragments by conflict-free 32-bit loads, fp32 softmax
// on the accumulator fragments (two query rows per lane), O = P.V with V fragments by ldmatrix.trans.
// attention_mma_kernel above runs one CTA per (image, head) and gathers 128-byte slices of 3 KB rows for K, V and
// Q separately: 2048 small CTAs per launch at config B against 256 here, each HBM byte requested once.
constexpr int ET_PITCH = 3072 + 16;

template <int NT>   // 8-key tiles: 7 (n <= 56) or 8 (n <= 64)
__global__ void __launch_bounds__(256)
encoder_self_attention_tc_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ key_mask, bf16* __restrict__ out,
                                 int ldo, int n, float scale) {
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t et_smem[];     // [n + 1][ET_PITCH] (last row zero), then NT*8 mask bytes
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    uint8_t* zero_row = et_smem + static_cast<size_t>(n) * ET_PITCH;
    uint8_t* smask = zero_row + ET_PITCH;
    const uint32_t bar_addr = xs_smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(static_cast<uint32_t>(n) * 3072u)
                     : "memory");
    }
    for (int i = threadIdx.x; i < ET_PITCH / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(zero_row)[i] = 0u;
    // Nothing a stream predecessor wrote may be read before this wait -- not even the key mask, which a kernel several
    // launches back produced: with every kernel triggering its dependents at entry, a whole run of kernels can be
    // resident before the first of them has finished (this read used to sit above the wait; two batches with different
    // masks back to back then gave run-to-run differences).
    pdl_wait();
    for (int j = threadIdx.x; j < NT * 8; j += blockDim.x) smask[j] = (j >= n || (key_mask && key_mask[static_cast<size_t>(b) * n + j])) ? 1 : 0;
    __syncthreads();   // barrier armed before any copy can complete on it; zero row and mask visible
    if (threadIdx.x < n) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(qkv) + (static_cast<size_t>(b) * n + threadIdx.x) * 3072;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         xs_smem_u32(et_smem + static_cast<size_t>(threadIdx.x) * ET_PITCH)),
                     "l"(reinterpret_cast<uint64_t>(src)), "r"(3072u), "r"(bar_addr)
                     : "memory");
    }
    cap_ptx::mbar_wait(&bar, 0);   // all rows landed (phase 0); bounded, leaves a fault record
    const int h = warp;
    const float sc = scale * 1.4426950408889634f;   // scores in the log2 domain
    const int mi = lane >> 3, mr = lane & 7;        // ldmatrix: this lane addresses row mr of 8x8 matrix mi
    for (int q0 = 0; q0 < n; q0 += 16) {
        const int ra = q0 + g, rb = ra + 8;
        const uint8_t* qa_row = (ra < n ? et_smem + static_cast<size_t>(ra) * ET_PITCH : zero_row) + h * (HEAD_DIM * 2) + 4 * t;
        const uint8_t* qb_row = (rb < n ? et_smem + static_cast<size_t>(rb) * ET_PITCH : zero_row) + h * (HEAD_DIM * 2) + 4 * t;
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(qa_row + ks * 32);
            qa[ks][1] = *reinterpret_cast<const uint32_t*>(qb_row + ks * 32);
            qa[ks][2] = *reinterpret_cast<const uint32_t*>(qa_row + ks * 32 + 16);
            qa[ks][3] = *reinterpret_cast<const uint32_t*>(qb_row + ks * 32 + 16);
        }
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            const int key = nt * 8 + g;
            const uint8_t* krow = (key < n ? et_smem + static_cast<size_t>(key) * ET_PITCH : zero_row) + 1024 + h * (HEAD_DIM * 2) + 4 * t;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                mma_bf16_16816(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(krow + ks * 32),
                               *reinterpret_cast<const uint32_t*>(krow + ks * 32 + 16));
        }
        float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x = smask[nt * 8 + 2 * t + (e & 1)] ? -INFINITY : s[nt][e] * sc;
                s[nt][e] = x;
                if (e < 2) ma = fmaxf(ma, x); else mb = fmaxf(mb, x);
            }
        }
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
        float suma = 0.f, sumb = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float m = e < 2 ? ma : mb;
                const float p = (m == -INFINITY) ? 0.f : fast_exp2(s[nt][e] - m);
                s[nt][e] = p;
                if (e < 2) suma += p; else sumb += p;
            }
        }
        suma += __shfl_xor_sync(0xffffffffu, suma, 1);
        suma += __shfl_xor_sync(0xffffffffu, suma, 2);
        sumb += __shfl_xor_sync(0xffffffffu, sumb, 1);
        sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);
        float o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < (NT + 1) / 2; ++kk) {
            constexpr int LAST = NT - 1;
            const int hi = 2 * kk + 1 <= LAST ? 2 * kk + 1 : LAST;   // clamp keeps the index constant-foldable
            const bool has_hi = 2 * kk + 1 <= LAST;
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = has_hi ? pack_bf16x2(s[hi][0], s[hi][1]) : 0u;
            pa[3] = has_hi ? pack_bf16x2(s[hi][2], s[hi][3]) : 0u;
            const int key = kk * 16 + (mi & 1) * 8 + mr;
            const uint8_t* vrow = (key < n ? et_smem + static_cast<size_t>(key) * ET_PITCH : zero_row) + 2048 + h * (HEAD_DIM * 2) +
                                  (mi >> 1) * 16;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, xs_smem_u32(vrow + dp * 32));
                mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
                mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
            }
        }
        const float inva = suma > 0.f ? 1.f / suma : 0.f;
        const float invb = sumb > 0.f ? 1.f / sumb : 0.f;
        bf16* oa = out + (static_cast<size_t>(b) * n + ra) * ldo + h * HEAD_DIM + 2 * t;
        bf16* ob = out + (static_cast<size_t>(b) * n + rb) * ldo + h * HEAD_DIM + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            if (ra < n) *reinterpret_cast<bf162*>(oa + dt * 8) = __floats2bfloat162_rn(o[dt][0] * inva, o[dt][1] * inva);
            if (rb < n) *reinterpret_cast<bf162*>(ob + dt * 8) = __floats2bfloat162_rn(o[dt][2] * invb, o[dt][3] * invb);
        }
    }
}

template <int NT>
int launch_encoder_tc(const AttnDev& a, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(a.nk + 1) * ET_PITCH + NT * 8;
    static cap_device_once smem_once;
    CAP_PROPAGATE(cap_opt_in_smem(smem_once, encoder_self_attention_tc_kernel<NT>, 220 * 1024));
    CAP_LAUNCH((encoder_self_attention_tc_kernel<NT>), a.B, 256, smem, stream, a.q, a.mask, a.out, a.ldo, a.nk, a.scale);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("encoder_self_attention_tc_kernel");
}

int launch_attention(const AttnDev& a, cudaStream_t stream) {
    const int nk_all = a.nk + a.n_mem;
    {   // whole-image encoder self-attention (plain attention over fused q|k|v rows, one CTA per image)
        const int hd = a.H * HEAD_DIM;
        if (a.H == 8 && a.geometry == nullptr && a.n_mem == 0 && a.nq == a.nk && a.nk <= 64 &&
            a.k == a.q + hd && a.v == a.q + 2 * hd && a.ldq == 3 * hd && a.ldk == 3 * hd && a.ldv == 3 * hd &&
            a.q_bs == static_cast<long long>(a.nk) * 3 * hd && a.o_bs == static_cast<long long>(a.nk) * a.ldo &&
            (a.mask == nullptr || (a.mask_qs == 0 && a.mask_bs == a.nk)) && a.ldo % 2 == 0 &&
            (reinterpret_cast<uintptr_t>(a.q) & 15) == 0) {
            return a.nk <= 56 ? launch_encoder_tc<7>(a, stream) : launch_encoder_tc<8>(a, stream);
        }
    }
    if (nk_all <= 128 && a.sentinel == nullptr) {  // tensor-core variant (the per-query sentinel lives in the CUDA-core kernel)
        dim3 grid((a.nq + MMA_Q_TILE - 1) / MMA_Q_TILE, a.H, a.B);
        if (nk_all <= 64)
            CAP_LAUNCH((attention_mma_kernel<8>), grid, 128, 0, stream, a);
        else
            CAP_LAUNCH((attention_mma_kernel<16>), grid, 128, 0, stream, a);
        g_cap_launches.fetch_add(1, std::memory_order_relaxed);
        return cap_check_launch("attention_mma_kernel");
    }
    const size_t smem = static_cast<size_t>(nk_all) * (K_STRIDE + HEAD_DIM) * 2 + ATT_WARPS * HEAD_DIM * 4 +
                        ATT_WARPS * MAX_KEYS * 4;
    static cap_device_once smem_once;
    CAP_PROPAGATE(cap_opt_in_smem(smem_once, attention_kernel, 64 * 1024));
    dim3 grid((a.nq + Q_TILE - 1) / Q_TILE, a.H, a.B);
    CAP_LAUNCH((attention_kernel), grid, ATT_THREADS, smem, stream, a);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("attention_kernel");
}

}  // namespace

extern "C" int cap_attention(const cap_attention_args* args, cap_stream_t stream) {
    CAP_REQUIRE(args != nullptr, "cap_attention: null args");
    const cap_attention_args& g = *args;
    CAP_REQUIRE(g.q && g.k && g.v && g.out, "cap_attention: null tensor");
    CAP_REQUIRE(g.B > 0 && g.H > 0 && g.nq > 0 && g.nk > 0, "cap_attention: empty problem");
    CAP_REQUIRE(g.B <= 65535 && g.H <= 65535, "cap_attention: B and H must be <= 65535");
    CAP_REQUIRE(g.n_mem >= 0 && g.nk + g.n_mem <= MAX_KEYS, "cap_attention: nk + n_mem = %d exceeds %d",
                g.nk + g.n_mem, MAX_KEYS);
    CAP_REQUIRE(g.n_mem == 0 || (g.mem_k && g.mem_v), "cap_attention: memory slots requested without tensors");
    CAP_REQUIRE(g.ldq % 2 == 0 && g.ldk % 2 == 0 && g.ldv % 2 == 0 && g.ldo % 2 == 0,
                "cap_attention: leading dimensions must be even");
    AttnDev a;
    a.q = static_cast<const bf16*>(g.q);
    a.k = static_cast<const bf16*>(g.k);
    a.v = static_cast<const bf16*>(g.v);
    a.out = static_cast<bf16*>(g.out);
    a.q_bs = g.q_bs; a.k_bs = g.k_bs; a.v_bs = g.v_bs; a.o_bs = g.o_bs;
    a.ldq = g.ldq; a.ldk = g.ldk; a.ldv = g.ldv; a.ldo = g.ldo;
    a.mask = g.mask; a.mask_bs = g.mask_bs; a.mask_qs = g.mask_qs;
    a.geometry = g.geometry;
    a.mem_k = static_cast<const bf16*>(g.mem_k);
    a.mem_v = static_cast<const bf16*>(g.mem_v);
    a.n_mem = g.n_mem; a.B = g.B; a.H = g.H; a.nq = g.nq; a.nk = g.nk;
    a.scale = g.scale;
    a.sentinel = static_cast<const bf16*>(g.sentinel); a.s_bs = g.s_bs; a.lds = g.lds;
    CAP_REQUIRE(g.sentinel == nullptr || g.lds % 2 == 0, "cap_attention: sentinel leading dimension must be even");
    return launch_attention(a, static_cast<cudaStream_t>(stream));
}

extern "C" int cap_decode_cross_attention(const void* q, int ldq, const void* kv, const uint8_t* key_mask, void* out,
                                          int ldo, int B, int beam, int n, int H, float scale, cap_stream_t stream) {
    CAP_REQUIRE(q && kv && out, "cap_decode_cross_attention: null pointer");
    CAP_REQUIRE(B > 0 && beam > 0 && n > 0 && n <= MAX_KEYS && H > 0, "cap_decode_cross_attention: bad shape");
    const int hd = H * HEAD_DIM;
    if (H == 8 && beam <= 8 && n <= 104 && ldq % 2 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0) {
        const bf16* qp = static_cast<const bf16*>(q);
        const bf16* kvp = static_cast<const bf16*>(kv);
        bf16* op = static_cast<bf16*>(out);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        if (n <= 56) return launch_cross_stream<7>(qp, ldq, kvp, key_mask, op, ldo, B, beam, n, scale, s);
        return launch_cross_stream<13>(qp, ldq, kvp, key_mask, op, ldo, B, beam, n, scale, s);
    }
    // other shapes: the general attention kernels with the image's beams as the queries
    AttnDev a;
    a.q = static_cast<const bf16*>(q);
    a.k = static_cast<const bf16*>(kv);
    a.v = static_cast<const bf16*>(kv) + hd;
    a.out = static_cast<bf16*>(out);
    a.q_bs = static_cast<long long>(beam) * ldq;
    a.k_bs = a.v_bs = static_cast<long long>(n) * 2 * hd;
    a.o_bs = static_cast<long long>(beam) * ldo;
    a.ldq = ldq; a.ldk = a.ldv = 2 * hd; a.ldo = ldo;
    a.mask = key_mask; a.mask_bs = n; a.mask_qs = 0;
    a.geometry = nullptr; a.mem_k = a.mem_v = nullptr; a.n_mem = 0;
    a.sentinel = nullptr; a.s_bs = 0; a.lds = 0;
    a.B = B; a.H = H; a.nq = beam; a.nk = n;
    a.scale = scale;
    return launch_attention(a, static_cast<cudaStream_t>(stream));
}

// Cross-attention of the same queries over several encoder levels in one launch (MeshedDecoderLayer, decoders.py:55-57):
// level i reads kv + i * kv_level_stride elements ([B][n][K|V]) and writes out + i * out_level_stride elements.
extern "C" int cap_decode_cross_attention_levels(const void* q, int ldq, const void* kv, size_t kv_level_stride,
                                                 const uint8_t* key_mask, void* out, int ldo, size_t out_level_stride, int B,
                                                 int beam, int n, int H, int levels, float scale, cap_stream_t stream) {
    CAP_REQUIRE(q && kv && out, "cap_decode_cross_attention_levels: null pointer");
    CAP_REQUIRE(B > 0 && beam > 0 && n > 0 && n <= MAX_KEYS && H > 0 && levels >= 1, "cap_decode_cross_attention_levels: bad shape");
    if (H == 8 && beam <= 8 && n <= 104 && ldq % 2 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0 &&
        (kv_level_stride * 2) % 16 == 0) {
        const bf16* qp = static_cast<const bf16*>(q);
        const bf16* kvp = static_cast<const bf16*>(kv);
        bf16* op = static_cast<bf16*>(out);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        if (n <= 56) return launch_cross_stream<7>(qp, ldq, kvp, key_mask, op, ldo, B, beam, n, scale, s, levels, kv_level_stride, out_level_stride);
        return launch_cross_stream<13>(qp, ldq, kvp, key_mask, op, ldo, B, beam, n, scale, s, levels, kv_level_stride, out_level_stride);
    }
    for (int i = 0; i < levels; ++i)   // other shapes: one launch per level
        CAP_PROPAGATE(cap_decode_cross_attention(q, ldq, static_cast<const bf16*>(kv) + i * kv_level_stride, key_mask,
                                                 static_cast<bf16*>(out) + i * out_level_stride, ldo, B, beam, n, H, scale, stream));
    return CAP_OK;
}

extern "C" int cap_decode_self_attention(const void* qkv, const int32_t* ancestry, const uint8_t* padflag, void* out,
                                         int ldo, int t, int R, int H, float scale, cap_stream_t stream) {
    CAP_REQUIRE(qkv && ancestry && padflag && out, "cap_decode_self_attention: null pointer");
    CAP_REQUIRE(t >= 0 && t < DEC_MAX_T, "cap_decode_self_attention: step %d outside [0,%d)", t, DEC_MAX_T);
    CAP_REQUIRE(R > 0 && H > 0 && ldo % 2 == 0, "cap_decode_self_attention: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bf16* cache = static_cast<const bf16*>(qkv);
    bf16* o = static_cast<bf16*>(out);
    const int wide_blocks = (R + DEC_WARPS - 1) / DEC_WARPS;
    if (H == 8 && ldo % 8 == 0) {
        CAP_LAUNCH((decode_self_attention_wide_kernel<16>), wide_blocks, DEC_WARPS * 32, 0, s, cache, ancestry, padflag, o, ldo, t, R, scale);
    } else {   // other head counts: one warp per (row, head)
        const int items = R * H;
        CAP_LAUNCH((decode_self_attention_kernel), (items + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, 0, s, cache, ancestry, padflag, o, ldo, t, R, H, scale);
    }
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("decode_self_attention_kernel");
}
