// T1 -- the XE training step: the kernels the backward pass, the loss and the
// optimizer need next to the forward kernels of the caption path.  Reference: trainers/vi_trainer.py:105-119 (the
// step), trainers/base_trainer.py:89-91, 114-117 (Adam(lr, betas=(0.9, 0.98)), NLLLoss(ignore_index=<pad>), Noam
// schedule); the operators differentiated are the ones of models/modules/{attentions,encoders,decoders,
// positionwise_feed_forward}.py that the forward kernels implement.
//
// Every GEMM of the backward pass (dX = dY.W, dW = dY^T.X) runs on the tcgen05 GEMM of gemm_tcgen05.cu, which
// multiplies K-major operands: this file supplies the bf16 transposes that put M on the contraction axis (with the
// bias gradient -- the column sums of dY -- folded into the same pass), the residual + LayerNorm pair that keeps what
// its backward needs, the attention backward, ReLU backward, embedding forward / backward, the fused log-softmax +
// NLL loss + logits gradient, and Adam on one flat parameter buffer with the bf16 shadow copy the GEMMs read.
// All HBM-bound except the attention backward (fp32 CUDA-core math on shared-memory tiles; n, T <= 128).
#include "cap_common.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>

extern std::atomic<long long> g_cap_launches;

namespace {

constexpr int TR_ROWS = 8;          // warps (= rows in flight) per CTA of the row kernels
constexpr int TR_MAX_CHUNKS = 4;    // 4 chunks x 32 lanes x 8 elements: d <= 1024

__device__ __forceinline__ void ld8(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float* f) { *reinterpret_cast<bf16x8*>(p) = pack8(f); }

// ------------------------------------------------------------------------------------------ residual + LayerNorm
// pre = a (+ res); out = LN(pre) * gamma + beta (+ pos[row % pos_rows]); rows flagged in zero_rows are zeroed in the
// outputs.  `pre` is kept for the backward pass (fp32), the outputs are the fp32 residual stream and the bf16 GEMM
// operand.  attentions.py:308-309, positionwise_feed_forward.py:26, encoders.py:20,36, decoders.py:26.
__global__ void __launch_bounds__(TR_ROWS * 32)
train_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ res, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float eps, const float* __restrict__ pos, int pos_rows,
                           const uint8_t* __restrict__ zero_rows, float* __restrict__ pre, float* __restrict__ out32,
                           bf16* __restrict__ out16, int rows, int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * TR_ROWS + warp;
    if (row >= rows) return;
    const int nchunks = d >> 3;
    const size_t base = static_cast<size_t>(row) * d;
    float v[TR_MAX_CHUNKS][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
            ld8(a + base + c * 8, v[i]);
            if (res != nullptr) {
                float r[8];
                ld8(res + base + c * 8, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] += r[j];
            }
            st8(pre + base + c * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = v[i][j] - mean;
                sq += t * t;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / d + eps);
    const bool zero = zero_rows != nullptr && zero_rows[row] != 0;
    const float* prow = pos ? pos + static_cast<size_t>(row % pos_rows) * d : nullptr;
#pragma unroll
    for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
            float g[8], b[8], o[8];
            ld8(gamma + c * 8, g);
            ld8(beta + c * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = zero ? 0.f : (v[i][j] - mean) * rstd * g[j] + b[j];
            if (prow != nullptr && !zero) {
                float pp[8];
                ld8(prow + c * 8, pp);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += pp[j];
            }
            st8(out32 + base + c * 8, o);
            st8(out16 + base + c * 8, o);
        }
    }
}

// Backward of the above.  dout = dout_a (+ dout_b) is the gradient of the (zeroed, position-shifted) output; rows in
// zero_rows receive no gradient.  With xhat = (pre - mean) * rstd and gy = dout * gamma:
//   dpre = rstd * (gy - mean(gy) - xhat * mean(gy * xhat)),  dgamma += sum_rows dout * xhat,  dbeta += sum_rows dout.
// dpre is the gradient of BOTH summands of pre (the GEMM output and the residual); it is written in fp32 (residual
// stream) and bf16 (operand of the GEMMs' backward).  A CTA walks rows with a grid stride; each warp keeps its lanes'
// dgamma / dbeta partial sums in registers, the CTA reduces them through shared memory and issues one atomicAdd per
// column.
__global__ void __launch_bounds__(TR_ROWS * 32)
train_layernorm_bwd_kernel(const float* __restrict__ dout_a, const float* __restrict__ dout_b, const float* __restrict__ pre,
                           const float* __restrict__ gamma, float eps, const uint8_t* __restrict__ zero_rows,
                           float* __restrict__ dpre32, bf16* __restrict__ dpre16, float* __restrict__ dgamma,
                           float* __restrict__ dbeta, int rows, int d) {
    pdl_prologue();
    extern __shared__ float red[];   // [TR_ROWS][2 * d]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = d >> 3;
    float ag[TR_MAX_CHUNKS][8], ab[TR_MAX_CHUNKS][8];
#pragma unroll
    for (int i = 0; i < TR_MAX_CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
    for (int row = blockIdx.x * TR_ROWS + warp; row < rows; row += gridDim.x * TR_ROWS) {
        const size_t base = static_cast<size_t>(row) * d;
        const bool zero = zero_rows != nullptr && zero_rows[row] != 0;
        float x[TR_MAX_CHUNKS][8], g[TR_MAX_CHUNKS][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunks) {
                ld8(pre + base + c * 8, x[i]);
                if (zero) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[i][j] = 0.f;
                } else {
                    ld8(dout_a + base + c * 8, g[i]);
                    if (dout_b != nullptr) {
                        float t[8];
                        ld8(dout_b + base + c * 8, t);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[i][j] += t[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += x[i][j];
            }
        }
        const float mean = warp_sum(sum) / d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x[i][j] -= mean;
                    sq += x[i][j] * x[i][j];
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / d + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunks) {
                float gm[8];
                ld8(gamma + c * 8, gm);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x[i][j] *= rstd;                       // xhat
                    ag[i][j] += g[i][j] * x[i][j];
                    ab[i][j] += g[i][j];
                    g[i][j] *= gm[j];                      // gy
                    s1 += g[i][j];
                    s2 += g[i][j] * x[i][j];
                }
            }
        }
        s1 = warp_sum(s1) / d;
        s2 = warp_sum(s2) / d;
#pragma unroll
        for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunks) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - s1 - x[i][j] * s2);
                st8(dpre32 + base + c * 8, o);
                st8(dpre16 + base + c * 8, o);
            }
        }
    }
    float* mine = red + static_cast<size_t>(warp) * 2 * d;
#pragma unroll
    for (int i = 0; i < TR_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
            st8(mine + c * 8, ag[i]);
            st8(mine + d + c * 8, ab[i]);
        }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < 2 * d; col += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TR_ROWS; ++w) s += red[static_cast<size_t>(w) * 2 * d + col];
        if (s != 0.f) atomicAdd(col < d ? dgamma + col : dbeta + (col - d), s);
    }
}

// ------------------------------------------------------------------------------------------ transpose (+ column sums)
// out[c][r] = in[r][c] for r < rows, and 0 for rows <= r < ldo (the contraction axis of the GEMM that follows is
// padded to a multiple of 8 elements with zeros); colsum[c] += sum_r in[r][c] (fp32) when requested: the bias
// gradient of a Linear is the column sum of its output gradient.
__global__ void __launch_bounds__(256)
transpose_colsum_kernel(const bf16* __restrict__ in, int ld, bf16* __restrict__ out, int ldo, float* __restrict__ colsum,
                        int rows, int cols) {
    pdl_prologue();
    __shared__ bf16 tile[32][33];
    __shared__ float part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        bf16 v = __float2bfloat16(0.f);
        if (r < rows && c < cols) v = in[static_cast<size_t>(r) * ld + c];
        tile[ty + 8 * k][tx] = v;
        s += __bfloat162float(v);
    }
    if (colsum != nullptr) part[ty][tx] = s;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (c < cols && r < ldo) out[static_cast<size_t>(c) * ldo + r] = tile[tx][ty + 8 * k];
    }
    if (colsum != nullptr && ty == 0 && c0 + tx < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][tx];
        if (t != 0.f) atomicAdd(colsum + c0 + tx, t);
    }
}

// dh *= (h > 0): backward of the ReLU fused into fc1's epilogue (positionwise_feed_forward.py:24)
__global__ void __launch_bounds__(256)
relu_bwd_kernel(bf16* __restrict__ dh, const bf16* __restrict__ h, size_t n8) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float a[8], b[8];
        unpack8(reinterpret_cast<const bf16x8*>(dh)[i], a);
        unpack8(reinterpret_cast<const bf16x8*>(h)[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = b[j] > 0.f ? a[j] : 0.f;
        reinterpret_cast<bf16x8*>(dh)[i] = pack8(a);
    }
}

// dst += src (fp32): the encoder output's gradient is the sum over the decoder layers' cross-attention K|V projections
__global__ void __launch_bounds__(256)
axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n4) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 a = reinterpret_cast<float4*>(dst)[i];
        const float4 b = reinterpret_cast<const float4*>(src)[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(dst)[i] = a;
    }
}

// ------------------------------------------------------------------------------------------ attention backward
// One CTA per (batch, head).  With P = softmax(q.k^T * scale + mask) (recomputed), O = P.v, D_i = dO_i . O_i:
//   dV = P^T.dO,   dS = P * (dO.v^T - D),   dQ = scale * dS.k,   dK = scale * dS^T.q        (attentions.py:51-55)
// q, k, v, dO tiles and P live in shared memory as fp32; the six small matrix products run on 4 x 4 register tiles
// (one 16-byte shared-memory load per 8 FMAs; the first version, one output per thread with two scalar loads per FMA,
// was shared-memory-bandwidth bound at 280 us per launch).  Rows are 272 bytes apart, so the 8 lanes of a load phase
// that read 8 consecutive rows hit 8 different 16-byte bank groups; sizes are padded to multiples of 4 with zero rows /
// masked columns.
constexpr int AB_THREADS = 256;
constexpr int AB_PITCH = 68;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& b) {
    acc.x = fmaf(s, b.x, acc.x); acc.y = fmaf(s, b.y, acc.y); acc.z = fmaf(s, b.z, acc.z); acc.w = fmaf(s, b.w, acc.w);
}
__device__ __forceinline__ void store_bf16x4(bf16* p, const float4& v) {
    bf162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

// acc[u][v] = A[4 * ti + u] . B[tj + tiles_n * v]   (rows of A and B contracted over their 64 columns)
__device__ __forceinline__ void tile_nt(const float* A, const float* B, int ti, int tj, int tiles_n, float (&acc)[4][4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 64; c += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = ld4(A + (4 * ti + u) * AB_PITCH + c);
#pragma unroll
        for (int v = 0; v < 4; ++v) b[v] = ld4(B + (tj + tiles_n * v) * AB_PITCH + c);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] += dot4(a[u], b[v]);
    }
}

// acc[u] = sum_j P[4 * ti + u][j] * B[j][c0 .. c0 + 3]
__device__ __forceinline__ void tile_nn(const float* P, int pp, const float* B, int ti, int c0, int n, float4 (&acc)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
        const float4 b = ld4(B + j * AB_PITCH + c0);
#pragma unroll
        for (int u = 0; u < 4; ++u) fma4(acc[u], P[(4 * ti + u) * pp + j], b);
    }
}

// acc[v] = sum_i P[i][4 * tj + v] * A[i][c0 .. c0 + 3]
__device__ __forceinline__ void tile_tn(const float* P, int pp, const float* A, int tj, int c0, int n, float4 (&acc)[4]) {
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
        const float4 p4 = ld4(P + i * pp + 4 * tj);
        const float4 a = ld4(A + i * AB_PITCH + c0);
        fma4(acc[0], p4.x, a); fma4(acc[1], p4.y, a); fma4(acc[2], p4.z, a); fma4(acc[3], p4.w, a);
    }
}

__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_kernel(cap_attention_args a, const bf16* __restrict__ d_out, bf16* __restrict__ dq, bf16* __restrict__ dk,
                     bf16* __restrict__ dv) {
    pdl_prologue();
    extern __shared__ __align__(16) float ab_smem[];
    const int nq = a.nq, nk = a.nk;
    const int nqp = (nq + 3) & ~3, nkp = (nk + 3) & ~3;   // padded to whole 4 x 4 tiles
    const int tiles_m = nqp >> 2, tiles_n = nkp >> 2;
    const int pp = nkp + 4;                                // row pitch of P (a multiple of 4 floats)
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    float* Qs = ab_smem;
    float* Ks = Qs + nqp * AB_PITCH;
    float* Vs = Ks + nkp * AB_PITCH;
    float* Gs = Vs + nkp * AB_PITCH;         // dO
    float* P = Gs + nqp * AB_PITCH;          // [nqp][pp]
    float* D = P + nqp * pp;                 // [nqp]
    const bf16* qg = static_cast<const bf16*>(a.q) + b * a.q_bs + h * 64;
    const bf16* kg = static_cast<const bf16*>(a.k) + b * a.k_bs + h * 64;
    const bf16* vg = static_cast<const bf16*>(a.v) + b * a.v_bs + h * 64;
    const bf16* gg = d_out + b * a.o_bs + h * 64;
    for (int idx = threadIdx.x; idx < nqp * 32; idx += AB_THREADS) {   // bf16 pairs; pad rows are zero
        const int r = idx >> 5, c = (idx & 31) * 2;
        float2 x = make_float2(0.f, 0.f), y = make_float2(0.f, 0.f);
        if (r < nq) {
            x = __bfloat1622float2(*reinterpret_cast<const bf162*>(qg + static_cast<size_t>(r) * a.ldq + c));
            y = __bfloat1622float2(*reinterpret_cast<const bf162*>(gg + static_cast<size_t>(r) * a.ldo + c));
        }
        *reinterpret_cast<float2*>(Qs + r * AB_PITCH + c) = x;
        *reinterpret_cast<float2*>(Gs + r * AB_PITCH + c) = y;
    }
    for (int idx = threadIdx.x; idx < nkp * 32; idx += AB_THREADS) {
        const int r = idx >> 5, c = (idx & 31) * 2;
        float2 x = make_float2(0.f, 0.f), y = make_float2(0.f, 0.f);
        if (r < nk) {
            x = __bfloat1622float2(*reinterpret_cast<const bf162*>(kg + static_cast<size_t>(r) * a.ldk + c));
            y = __bfloat1622float2(*reinterpret_cast<const bf162*>(vg + static_cast<size_t>(r) * a.ldv + c));
        }
        *reinterpret_cast<float2*>(Ks + r * AB_PITCH + c) = x;
        *reinterpret_cast<float2*>(Vs + r * AB_PITCH + c) = y;
    }
    __syncthreads();
    // S = q.k^T * scale; masked entries and pad columns -inf
    const uint8_t* mask = a.mask ? a.mask + b * a.mask_bs : nullptr;
    for (int t = threadIdx.x; t < tiles_m * tiles_n; t += AB_THREADS) {
        const int ti = t / tiles_n, tj = t - ti * tiles_n;
        float acc[4][4];
        tile_nt(Qs, Ks, ti, tj, tiles_n, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * ti + u;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int j = tj + tiles_n * v;
                float sv = acc[u][v] * a.scale;
                if (j >= nk || (mask != nullptr && i < nq && mask[static_cast<size_t>(i) * a.mask_qs + j])) sv = -INFINITY;
                P[i * pp + j] = sv;
            }
        }
    }
    __syncthreads();
    // row softmax (one warp per row)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < nqp; i += AB_THREADS / 32) {
        float m = -INFINITY;
        for (int j = lane; j < nkp; j += 32) m = fmaxf(m, P[i * pp + j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < nkp; j += 32) {
            const float e = (m == -INFINITY) ? 0.f : __expf(P[i * pp + j] - m);
            P[i * pp + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        for (int j = lane; j < nkp; j += 32) P[i * pp + j] *= inv;
    }
    __syncthreads();
    // D_i = dO_i . O_i with O = P.v: 4 rows x 4 columns per thread, the 16 column groups of a row tile are 16 lanes
    for (int t0 = 0; t0 < tiles_m * 16; t0 += AB_THREADS) {
        const int t = t0 + threadIdx.x;
        const bool active = t < tiles_m * 16;
        const int ti = active ? t >> 4 : 0, c0 = (t & 15) * 4;
        float4 acc[4];
        tile_nn(P, pp, Vs, ti, c0, nkp, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float dsum = active ? dot4(acc[u], ld4(Gs + (4 * ti + u) * AB_PITCH + c0)) : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
            if (active && (t & 15) == 0) D[4 * ti + u] = dsum;
        }
    }
    // dV[j][c] = sum_i P[i][j] * dO[i][c]
    bf16* dvg = dv + b * a.v_bs + h * 64;
    for (int t = threadIdx.x; t < tiles_n * 16; t += AB_THREADS) {
        const int tj = t >> 4, c0 = (t & 15) * 4;
        float4 acc[4];
        tile_tn(P, pp, Gs, tj, c0, nqp, acc);
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (4 * tj + v < nk) store_bf16x4(dvg + static_cast<size_t>(4 * tj + v) * a.ldv + c0, acc[v]);
    }
    __syncthreads();
    // dS in place: P * (dO.v^T - D) * scale
    for (int t = threadIdx.x; t < tiles_m * tiles_n; t += AB_THREADS) {
        const int ti = t / tiles_n, tj = t - ti * tiles_n;
        float acc[4][4];
        tile_nt(Gs, Vs, ti, tj, tiles_n, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * ti + u;
            const float di = D[i];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int j = tj + tiles_n * v;
                P[i * pp + j] = P[i * pp + j] * (acc[u][v] - di) * a.scale;
            }
        }
    }
    __syncthreads();
    bf16* dqg = dq + b * a.q_bs + h * 64;
    for (int t = threadIdx.x; t < tiles_m * 16; t += AB_THREADS) {
        const int ti = t >> 4, c0 = (t & 15) * 4;
        float4 acc[4];
        tile_nn(P, pp, Ks, ti, c0, nkp, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (4 * ti + u < nq) store_bf16x4(dqg + static_cast<size_t>(4 * ti + u) * a.ldq + c0, acc[u]);
    }
    bf16* dkg = dk + b * a.k_bs + h * 64;
    for (int t = threadIdx.x; t < tiles_n * 16; t += AB_THREADS) {
        const int tj = t >> 4, c0 = (t & 15) * 4;
        float4 acc[4];
        tile_tn(P, pp, Qs, tj, c0, nqp, acc);
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (4 * tj + v < nk) store_bf16x4(dkg + static_cast<size_t>(4 * tj + v) * a.ldk + c0, acc[v]);
    }
}

// ------------------------------------------------------------------------------------------ embedding
// out[row] = emb[token] + pos[token == pad ? 0 : row % T + 1]  (decoders.py:105-112, non-stateful branch)
__global__ void __launch_bounds__(TR_ROWS * 32)
train_embed_fwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ pos,
                       int T, int pad_idx, float* __restrict__ out32, bf16* __restrict__ out16, int rows, int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * TR_ROWS + warp;
    if (row >= rows) return;
    const int64_t tok = tokens[row];
    const int p = tok == pad_idx ? 0 : row % T + 1;
    const float* e = emb + static_cast<size_t>(tok) * d;
    const float* q = pos + static_cast<size_t>(p) * d;
    for (int c = lane; c < (d >> 3); c += 32) {
        float x[8], y[8];
        ld8(e + c * 8, x);
        ld8(q + c * 8, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        st8(out32 + static_cast<size_t>(row) * d + c * 8, x);
        st8(out16 + static_cast<size_t>(row) * d + c * 8, x);
    }
}

// d_emb[token] += g_a[row] (+ g_b[row]); the <pad> row never receives a gradient (nn.Embedding(padding_idx))
__global__ void __launch_bounds__(TR_ROWS * 32)
train_embed_bwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ g_a, const float* __restrict__ g_b,
                       int pad_idx, float* __restrict__ d_emb, int rows, int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * TR_ROWS + warp;
    if (row >= rows) return;
    const int64_t tok = tokens[row];
    if (tok == pad_idx) return;
    float* dst = d_emb + static_cast<size_t>(tok) * d;
    for (int c = lane; c < d; c += 32) {
        float g = g_a[static_cast<size_t>(row) * d + c];
        if (g_b != nullptr) g += g_b[static_cast<size_t>(row) * d + c];
        atomicAdd(dst + c, g);
    }
}

// ------------------------------------------------------------------------------------------ loss
// stats[0] += number of targets != ignore
__global__ void __launch_bounds__(256)
count_targets_kernel(const int64_t* __restrict__ targets, int ignore, float* __restrict__ stats, int rows) {
    pdl_prologue();
    int n = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) n += targets[i] != ignore;
    n = static_cast<int>(warp_sum(static_cast<float>(n)));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(stats, static_cast<float>(n));
}

// Per row: lse = log sum exp(logits); stats[1] += lse - logits[target]; dlogits = (softmax - onehot) / stats[0], zero for
// ignored rows and for the padding columns V .. ld.  NLLLoss(ignore_index) over log_softmax (base_trainer.py:91,
// decoders.py:123), mean over the counted targets.  With row_weight: stats[1] += w[row] * nll, dlogits = w[row] *
// (softmax - onehot) -- the self-critical loss -(mean_T log p) * (r - mean_b r), mean over B * b rows
// (vi_trainer.py:146-148), whose per-token weight is advantage / (T * B * b); finished positions carry <pad> targets.
__global__ void __launch_bounds__(256)
xent_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ targets, int ignore,
            const float* __restrict__ row_weight, float* __restrict__ stats, bf16* __restrict__ dlogits, int ldd, int V) {
    pdl_prologue();
    __shared__ float sh[8];
    const int row = blockIdx.x;
    const float* x = logits + static_cast<size_t>(row) * ld;
    bf16* g = dlogits + static_cast<size_t>(row) * ldd;
    const int64_t tgt = targets[row];
    if (tgt == ignore) {
        for (int c = threadIdx.x; c < ldd; c += blockDim.x) g[c] = __float2bfloat16(0.f);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < V; c += blockDim.x) m = fmaxf(m, x[c]);
    m = warp_max(m);
    if (lane == 0) sh[warp] = m;
    __syncthreads();
    m = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) s += __expf(x[c] - m);
    s = warp_sum(s);
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w];
    const float lse = m + logf(s);
    // mean over the counted targets (XE), or the caller's per-row weight (the self-critical loss: advantage / (T B b))
    const float inv_n = row_weight != nullptr ? row_weight[row] : 1.f / stats[0];
    if (threadIdx.x == 0) atomicAdd(stats + 1, row_weight != nullptr ? (lse - x[tgt]) * inv_n : lse - x[tgt]);
    for (int c = threadIdx.x; c < ldd; c += blockDim.x) {
        float v = 0.f;
        if (c < V) v = (__expf(x[c] - lse) - (c == tgt ? 1.f : 0.f)) * inv_n;
        g[c] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (no weight decay, no amsgrad) on one flat fp32 buffer, plus the bf16 shadow copy the GEMMs read:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with step_size = lr_t / (1 - b1^t) computed by the caller (lr_t carries the Noam factor, base_trainer.py:114-117).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ shadow, size_t n, float b1, float b2, float step_size, float inv_bc2_sqrt, float eps) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pi = p[i] - step_size * mi / (sqrtf(vi) * inv_bc2_sqrt + eps);
        p[i] = pi;
        shadow[i] = __float2bfloat16(pi);
    }
}

// ------------------------------------------------------------------------------------------ dropout
// x[i] = keep(i) ? x[i] * scale : 0, in place, with keep(i) = hash(i, seed, site) >= threshold: a counter-based mask,
// so the backward pass regenerates it from (seed, site) instead of storing it, and the oracle (and the patched
// reference of oracle/ref_harness/gen_golden_train.py) can generate the very same mask on the CPU.  The reference's
// sites are nn.Dropout modules (vision_embeddings.py:18, attentions.py:308, positionwise_feed_forward.py:24-25); `site`
// is the CRC-32 of the module's qualified name.
__device__ __forceinline__ uint32_t dropout_hash(uint32_t idx, uint32_t seed, uint32_t site) {
    uint32_t x = idx * 0x9E3779B1u + seed * 0x85EBCA77u + site * 0xC2B2AE3Du;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
    return x;
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(T* __restrict__ x, size_t n, uint32_t threshold, float scale, uint32_t seed, uint32_t site) {
    pdl_prologue();
    constexpr int VEC = 16 / sizeof(T);     // one 16-byte vector per thread and iteration; n is a multiple of VEC
    const size_t nv = n / VEC;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nv; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint4 raw = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t base = static_cast<uint32_t>(i * VEC);
        if constexpr (sizeof(T) == 2) {
            bf16* e = reinterpret_cast<bf16*>(&raw);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                e[j] = dropout_hash(base + j, seed, site) >= threshold ? __float2bfloat16(__bfloat162float(e[j]) * scale) : __float2bfloat16(0.f);
        } else {
            float* e = reinterpret_cast<float*>(&raw);
#pragma unroll
            for (int j = 0; j < VEC; ++j) e[j] = dropout_hash(base + j, seed, site) >= threshold ? e[j] * scale : 0.f;
        }
        reinterpret_cast<uint4*>(x)[i] = raw;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_tail_kernel(T* __restrict__ x, size_t begin, size_t n, uint32_t threshold, float scale, uint32_t seed, uint32_t site) {
    pdl_prologue();
    const size_t i = begin + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const bool keep = dropout_hash(static_cast<uint32_t>(i), seed, site) >= threshold;
    float v = 0.f;
    if (keep) {
        if constexpr (sizeof(T) == 2) v = __bfloat162float(x[i]) * scale; else v = x[i] * scale;
    }
    if constexpr (sizeof(T) == 2) x[i] = __float2bfloat16(v); else x[i] = v;
}

// out = sum over z of parts[z] (fp32): the partial products of a split-K GEMM (cap_linear_splitk)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ parts, int splits, size_t stride4, float* __restrict__ out, size_t n4) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 acc = reinterpret_cast<const float4*>(parts)[i];
        for (int z = 1; z < splits; ++z) {
            const float4 b = reinterpret_cast<const float4*>(parts)[z * stride4 + i];
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
}

inline int grid_for(size_t items, int block) {
    const size_t g = (items + block - 1) / block;
    return static_cast<int>(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int cap_train_layernorm_fwd(const float* a, const float* res, const float* gamma, const float* beta, float eps,
                                       const float* pos, int pos_rows, const uint8_t* zero_rows, float* pre, float* out_f32,
                                       void* out_bf16, int rows, int d, cap_stream_t stream) {
    CAP_REQUIRE(a && gamma && beta && pre && out_f32 && out_bf16, "cap_train_layernorm_fwd: null pointer");
    CAP_REQUIRE(rows > 0 && d % 8 == 0 && d <= TR_MAX_CHUNKS * 256, "cap_train_layernorm_fwd: d must be a multiple of 8, <= 1024");
    CAP_REQUIRE(!pos || pos_rows > 0, "cap_train_layernorm_fwd: pos_rows");
    CAP_LAUNCH(train_layernorm_fwd_kernel, (rows + TR_ROWS - 1) / TR_ROWS, TR_ROWS * 32, 0, static_cast<cudaStream_t>(stream), a, res,
               gamma, beta, eps, pos, pos_rows, zero_rows, pre, out_f32, static_cast<bf16*>(out_bf16), rows, d);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("train_layernorm_fwd_kernel");
}

extern "C" int cap_train_layernorm_bwd(const float* dout_a, const float* dout_b, const float* pre, const float* gamma, float eps,
                                       const uint8_t* zero_rows, float* dpre_f32, void* dpre_bf16, float* dgamma, float* dbeta,
                                       int rows, int d, cap_stream_t stream) {
    CAP_REQUIRE(dout_a && pre && gamma && dpre_f32 && dpre_bf16 && dgamma && dbeta, "cap_train_layernorm_bwd: null pointer");
    CAP_REQUIRE(rows > 0 && d % 8 == 0 && d <= TR_MAX_CHUNKS * 256, "cap_train_layernorm_bwd: d must be a multiple of 8, <= 1024");
    const int grid = std::min((rows + TR_ROWS - 1) / TR_ROWS, 148 * 2);
    const size_t smem = static_cast<size_t>(TR_ROWS) * 2 * d * sizeof(float);
    static cap_device_once once;
    CAP_PROPAGATE(cap_opt_in_smem(once, train_layernorm_bwd_kernel, 64 * 1024));
    CAP_LAUNCH(train_layernorm_bwd_kernel, grid, TR_ROWS * 32, smem, static_cast<cudaStream_t>(stream), dout_a, dout_b, pre, gamma, eps,
               zero_rows, dpre_f32, static_cast<bf16*>(dpre_bf16), dgamma, dbeta, rows, d);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("train_layernorm_bwd_kernel");
}

extern "C" int cap_transpose_bf16(const void* in, int ld, void* out, int ldo, float* colsum, int rows, int cols,
                                  cap_stream_t stream) {
    CAP_REQUIRE(in && out, "cap_transpose_bf16: null pointer");
    CAP_REQUIRE(rows > 0 && cols > 0 && ld >= cols && ldo >= rows, "cap_transpose_bf16: bad shape rows=%d cols=%d ld=%d ldo=%d", rows, cols, ld, ldo);
    dim3 grid((cols + 31) / 32, (ldo + 31) / 32);
    CAP_LAUNCH(transpose_colsum_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(in), ld,
               static_cast<bf16*>(out), ldo, colsum, rows, cols);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("transpose_colsum_kernel");
}

extern "C" int cap_train_relu_bwd(void* dh, const void* h, int64_t count, cap_stream_t stream) {
    CAP_REQUIRE(dh && h && count > 0 && count % 8 == 0, "cap_train_relu_bwd: count must be a positive multiple of 8");
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(dh) & 15) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0, "cap_train_relu_bwd: 16-byte alignment");
    const size_t n8 = static_cast<size_t>(count) / 8;
    CAP_LAUNCH(relu_bwd_kernel, grid_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream), static_cast<bf16*>(dh),
               static_cast<const bf16*>(h), n8);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("relu_bwd_kernel");
}

extern "C" int cap_axpy_f32(float* dst, const float* src, int64_t count, cap_stream_t stream) {
    CAP_REQUIRE(dst && src && count > 0 && count % 4 == 0, "cap_axpy_f32: count must be a positive multiple of 4");
    const size_t n4 = static_cast<size_t>(count) / 4;
    CAP_LAUNCH(axpy_kernel, grid_for(n4, 256), 256, 0, static_cast<cudaStream_t>(stream), dst, src, n4);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("axpy_kernel");
}

extern "C" int cap_attention_backward(const cap_attention_args* a, const void* d_out, void* dq, void* dk, void* dv,
                                      cap_stream_t stream) {
    CAP_REQUIRE(a && a->q && a->k && a->v && d_out && dq && dk && dv, "cap_attention_backward: null pointer");
    CAP_REQUIRE(!a->geometry && !a->mem_k && !a->mem_v && !a->sentinel,
                "cap_attention_backward: only the plain scaled dot-product attention (attentions.py:44-58) is differentiated");
    CAP_REQUIRE(a->nq >= 1 && a->nk >= 1 && a->nq <= 128 && a->nk <= 128, "cap_attention_backward: 1..128 queries and keys");
    CAP_REQUIRE(a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && a->ldo % 2 == 0 && a->q_bs % 4 == 0 && a->k_bs % 4 == 0 &&
                    a->v_bs % 4 == 0 && a->o_bs % 2 == 0,
                "cap_attention_backward: row / batch strides of q, k, v must be multiples of 4 elements (8-byte gradient stores)");
    CAP_REQUIRE((reinterpret_cast<uintptr_t>(dq) & 7) == 0 && (reinterpret_cast<uintptr_t>(dk) & 7) == 0 && (reinterpret_cast<uintptr_t>(dv) & 7) == 0,
                "cap_attention_backward: dq / dk / dv must be 8-byte aligned");
    const int nqp = (a->nq + 3) & ~3, nkp = (a->nk + 3) & ~3;
    const size_t smem = (static_cast<size_t>(2 * nqp + 2 * nkp) * AB_PITCH + static_cast<size_t>(nqp) * (nkp + 4) + nqp) * sizeof(float);
    static cap_device_once once;
    CAP_PROPAGATE(cap_opt_in_smem(once, attention_bwd_kernel, 227 * 1024));
    CAP_REQUIRE(smem <= 227 * 1024, "cap_attention_backward: tiles do not fit shared memory");
    CAP_LAUNCH(attention_bwd_kernel, a->B * a->H, AB_THREADS, smem, static_cast<cudaStream_t>(stream), *a,
               static_cast<const bf16*>(d_out), static_cast<bf16*>(dq), static_cast<bf16*>(dk), static_cast<bf16*>(dv));
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("attention_bwd_kernel");
}

extern "C" int cap_train_embed_fwd(const int64_t* tokens, const float* emb, const float* pos, int T, int pad_idx, float* out_f32,
                                   void* out_bf16, int rows, int d, cap_stream_t stream) {
    CAP_REQUIRE(tokens && emb && pos && out_f32 && out_bf16 && rows > 0 && T > 0 && d % 8 == 0, "cap_train_embed_fwd: bad arguments");
    CAP_LAUNCH(train_embed_fwd_kernel, (rows + TR_ROWS - 1) / TR_ROWS, TR_ROWS * 32, 0, static_cast<cudaStream_t>(stream), tokens, emb,
               pos, T, pad_idx, out_f32, static_cast<bf16*>(out_bf16), rows, d);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("train_embed_fwd_kernel");
}

extern "C" int cap_train_embed_bwd(const int64_t* tokens, const float* g_a, const float* g_b, int pad_idx, float* d_emb, int rows,
                                   int d, cap_stream_t stream) {
    CAP_REQUIRE(tokens && g_a && d_emb && rows > 0 && d > 0, "cap_train_embed_bwd: bad arguments");
    CAP_LAUNCH(train_embed_bwd_kernel, (rows + TR_ROWS - 1) / TR_ROWS, TR_ROWS * 32, 0, static_cast<cudaStream_t>(stream), tokens, g_a,
               g_b, pad_idx, d_emb, rows, d);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("train_embed_bwd_kernel");
}

extern "C" int cap_train_xent(const float* logits, int ld, const int64_t* targets, int ignore_index, const float* row_weight,
                              float* stats, void* dlogits, int ldd, int rows, int V, cap_stream_t stream) {
    CAP_REQUIRE(logits && targets && stats && dlogits, "cap_train_xent: null pointer");
    CAP_REQUIRE(rows > 0 && V > 0 && ld >= V && ldd >= V, "cap_train_xent: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CAP_CHECK_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(float), s));
    CAP_LAUNCH_SERIAL(count_targets_kernel, grid_for(rows, 256), 256, 0, s, targets, ignore_index, stats, rows);
    CAP_LAUNCH_SERIAL(xent_kernel, rows, 256, 0, s, logits, ld, targets, ignore_index, row_weight, stats, static_cast<bf16*>(dlogits),
                      ldd, V);
    g_cap_launches.fetch_add(2, std::memory_order_relaxed);
    return cap_check_launch("xent_kernel");
}

extern "C" int cap_train_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                              int64_t count, float lr, float beta1, float beta2, float eps, int step, cap_stream_t stream) {
    CAP_REQUIRE(params && grads && exp_avg && exp_avg_sq && shadow_bf16 && count > 0 && step >= 1, "cap_train_adam: bad arguments");
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    const float step_size = static_cast<float>(lr / bc1);
    const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
    CAP_LAUNCH_SERIAL(adam_kernel, grid_for(static_cast<size_t>(count), 256), 256, 0, static_cast<cudaStream_t>(stream), params, grads,
                      exp_avg, exp_avg_sq, static_cast<bf16*>(shadow_bf16), static_cast<size_t>(count), beta1, beta2, step_size,
                      inv_bc2_sqrt, eps);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("adam_kernel");
}

extern "C" int cap_train_dropout(void* x, int dtype, int64_t count, unsigned int threshold, float scale, unsigned int seed,
                                 unsigned int site, cap_stream_t stream) {
    CAP_REQUIRE(x && count > 0 && count <= 0xFFFFFFFFll, "cap_train_dropout: 1 .. 2^32 - 1 elements");
    CAP_REQUIRE(dtype == CAP_BF16 || dtype == CAP_F32, "cap_train_dropout: bf16 or fp32");
    const size_t n = static_cast<size_t>(count);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (dtype == CAP_BF16) {
        const size_t body = aligned ? n / 8 * 8 : 0;
        if (body) CAP_LAUNCH(dropout_kernel<bf16>, grid_for(body / 8, 256), 256, 0, s, static_cast<bf16*>(x), body, threshold, scale, seed, site);
        if (body < n) CAP_LAUNCH(dropout_tail_kernel<bf16>, static_cast<int>((n - body + 255) / 256), 256, 0, s, static_cast<bf16*>(x), body, n, threshold, scale, seed, site);
    } else {
        const size_t body = aligned ? n / 4 * 4 : 0;
        if (body) CAP_LAUNCH(dropout_kernel<float>, grid_for(body / 4, 256), 256, 0, s, static_cast<float*>(x), body, threshold, scale, seed, site);
        if (body < n) CAP_LAUNCH(dropout_tail_kernel<float>, static_cast<int>((n - body + 255) / 256), 256, 0, s, static_cast<float*>(x), body, n, threshold, scale, seed, site);
    }
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("dropout_kernel");
}

extern "C" int cap_sum_partials(const float* partials, int splits, int64_t count, float* out, cap_stream_t stream) {
    CAP_REQUIRE(partials && out && splits >= 1 && count > 0 && count % 4 == 0, "cap_sum_partials: count must be a positive multiple of 4");
    const size_t n4 = static_cast<size_t>(count) / 4;
    CAP_LAUNCH(sum_partials_kernel, grid_for(n4, 256), 256, 0, static_cast<cudaStream_t>(stream), partials, splits, n4, out, n4);
    g_cap_launches.fetch_add(1, std::memory_order_relaxed);
    return cap_check_launch("sum_partials_kernel");
}
