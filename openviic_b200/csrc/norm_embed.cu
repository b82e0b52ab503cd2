// HBM-bound row kernels of the caption path: residual+LayerNorm, visual padding mask + cast,
// box-relation bias, token embedding, meshed gate mix, AoA gate.  One warp per row, 128-bit
// vectorised accesses, warp-shuffle reductions, fp32 statistics.
#include "cap_common.cuh"

#include <atomic>

extern std::atomic<long long> g_cap_launches;

namespace {

constexpr int LN_MAX_CHUNKS = 8;  // 8 chunks * 32 lanes * 8 elements = d <= 2048
constexpr int ROWS_PER_BLOCK = 8; // 8 warps

__device__ __forceinline__ void load8(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ void load8(const bf16* p, float* f) {
    const bf16x8 v = *reinterpret_cast<const bf16x8*>(p);
    unpack8(v, f);
}

// out = LN(residual + y) * gamma + beta (+ pos) ; optionally zero whole rows.
__device__ __forceinline__ void store8(bf16* p, const float* f) { *reinterpret_cast<bf16x8*>(p) = pack8(f); }
__device__ __forceinline__ void store8(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Outputs: a bf16 copy (the next GEMM's A operand) and/or an fp32 copy (the next residual): the
// residual stream stays fp32 end to end, only GEMM inputs are rounded to bf16.
template <typename YT, typename RT>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
add_layernorm_kernel(const YT* __restrict__ y, int ldy, const RT* __restrict__ res, int ldr,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     const float* __restrict__ pos, int pos_rows, const uint8_t* __restrict__ zero_rows,
                     bf16* __restrict__ out, int ldo, float* __restrict__ out32, int ldo32, int rows, int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_BLOCK + warp;
    if (row >= rows) return;
    const int nchunks = d >> 3;
    bf16* orow = out ? out + static_cast<size_t>(row) * ldo : nullptr;
    float* orow32 = out32 ? out32 + static_cast<size_t>(row) * ldo32 : nullptr;
    if (zero_rows != nullptr && zero_rows[row]) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < nchunks; c += 32) {
            if (orow) store8(orow + c * 8, z);
            if (orow32) store8(orow32 + c * 8, z);
        }
        return;
    }
    float v[LN_MAX_CHUNKS][8];
    float sum = 0.f;
    const YT* yrow = y + static_cast<size_t>(row) * ldy;
    const RT* rrow = res ? res + static_cast<size_t>(row) * ldr : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
            load8(yrow + c * 8, v[i]);
            if (rrow) {
                float r[8];
                load8(rrow + c * 8, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] += r[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
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
    const float* prow = pos ? pos + static_cast<size_t>(row % pos_rows) * d : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
            float g[8], b[8], o[8];
            load8(gamma + c * 8, g);
            load8(beta + c * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
            if (prow) {
                float pp[8];
                load8(prow + c * 8, pp);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += pp[j];
            }
            if (orow) store8(orow + c * 8, o);
            if (orow32) store8(orow32 + c * 8, o);
        }
    }
}

// mask[row] = (fp32 sum of the raw feature row == 0); out = bf16(feats)
template <typename FT>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
feature_mask_cast_kernel(const FT* __restrict__ feats, bf16* __restrict__ out, uint8_t* __restrict__ mask, int rows,
                         int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_BLOCK + warp;
    if (row >= rows) return;
    const FT* src = feats + static_cast<size_t>(row) * d;
    bf16* dst = out ? out + static_cast<size_t>(row) * d : nullptr;
    float sum = 0.f;
    for (int c = lane; c < (d >> 3); c += 32) {
        float f[8];
        load8(src + c * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[j];
        if (dst) *reinterpret_cast<bf16x8*>(dst + c * 8) = pack8(f);
    }
    sum = warp_sum(sum);
    if (lane == 0) mask[row] = (sum == 0.f) ? 1 : 0;
}

// g[b,h,i,j] = relu(W_g[h,:] . emb(box_i, box_j) + b_g[h])
__global__ void geometry_bias_kernel(const float* __restrict__ boxes, const float* __restrict__ w_g,
                                     const float* __restrict__ b_g, float* __restrict__ g, int B, int n, int H,
                                     int d_g, int trig) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = B * n * n;
    if (idx >= total) return;
    const int j = idx % n, i = (idx / n) % n, b = idx / (n * n);
    const float4 bi = *reinterpret_cast<const float4*>(boxes + (static_cast<size_t>(b) * n + i) * 4);
    const float4 bj = *reinterpret_cast<const float4*>(boxes + (static_cast<size_t>(b) * n + j) * 4);
    const float cxi = (bi.x + bi.z) * 0.5f, cyi = (bi.y + bi.w) * 0.5f;
    const float wi = (bi.z - bi.x) + 1.f, hi = (bi.w - bi.y) + 1.f;
    const float cxj = (bj.x + bj.z) * 0.5f, cyj = (bj.y + bj.w) * 0.5f;
    const float wj = (bj.z - bj.x) + 1.f, hj = (bj.w - bj.y) + 1.f;
    float delta[4];
    delta[0] = logf(fmaxf(fabsf((cxi - cxj) / wi), 1e-3f));
    delta[1] = logf(fmaxf(fabsf((cyi - cyj) / hi), 1e-3f));
    delta[2] = logf(wi / wj);
    delta[3] = logf(hi / hj);
    for (int h = 0; h < H; ++h) {
        const float* wh = w_g + static_cast<size_t>(h) * d_g;
        float acc = b_g[h];
        if (!trig) {
#pragma unroll
            for (int c = 0; c < 4; ++c) acc += wh[c] * delta[c];
        } else {
            const int nf = d_g / 8;  // frequencies per coordinate
            for (int c = 0; c < 4; ++c) {
                for (int k = 0; k < nf; ++k) {
                    const float freq = 1.f / powf(1000.f, static_cast<float>(k) / static_cast<float>(nf));
                    const float a = 100.f * delta[c] * freq;
                    acc += wh[c * nf + k] * sinf(a) + wh[4 * nf + c * nf + k] * cosf(a);
                }
            }
        }
        g[((static_cast<size_t>(b) * H + h) * n + i) * n + j] = fmaxf(acc, 0.f);
    }
}

// Locally-constrained attention mask of the dual-path encoder (models/utils.py:100-154, get_combine_masks): for box i
// the grid cells inside its corner-to-corner cell rectangle stay visible (0), every other cell is masked (1).
// lower_bound(edges, v) = last k with k / g <= v (0 when none), compared in double like numpy's float64 edges.
__global__ void region_grid_mask_kernel(const float* __restrict__ boxes, uint8_t* __restrict__ mask, int rows, int g) {
    pdl_prologue();
    const int cells = g * g;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cells) return;
    const int cell = idx % cells, r = idx / cells;
    const float4 bx = *reinterpret_cast<const float4*>(boxes + static_cast<size_t>(r) * 4);
    auto last_le = [g](float v) {
        int pos = 0;
        for (int k = 0; k < g; ++k)
            if (static_cast<double>(k) / static_cast<double>(g) <= static_cast<double>(v)) pos = k;
        return pos;
    };
    const int x1 = last_le(bx.x), y1 = last_le(bx.y), x2 = last_le(bx.z), y3 = last_le(bx.w);
    const int top_left = y1 * g + x1, bot_left = y3 * g + x1, width = x2 - x1 + 1;
    const int rel = cell - top_left;
    bool covered = false;
    if (rel >= 0) {
        const int row = rel / g, col = rel % g;
        covered = (row * g + top_left <= bot_left) && (col < width);
    }
    mask[idx] = covered ? 0 : 1;
}

// x[r] = emb[token[r]] + pos_table[position]; padflag[r] = token == pad
__global__ void embed_tokens_kernel(const int32_t* __restrict__ tokens, const bf16* __restrict__ emb,
                                    const float* __restrict__ pos_table, int position, int pad_idx,
                                    bf16* __restrict__ out, float* __restrict__ out32,
                                    uint8_t* __restrict__ padflag, int R, int d) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_BLOCK + warp;
    if (row >= R) return;
    const int tok = tokens[row];
    if (lane == 0 && padflag) padflag[row] = (tok == pad_idx) ? 1 : 0;
    const bf16* e = emb + static_cast<size_t>(tok) * d;
    const float* p = pos_table + static_cast<size_t>(position) * d;
    bf16* o = out + static_cast<size_t>(row) * d;
    for (int c = lane; c < (d >> 3); c += 32) {
        float a[8], b[8];
        load8(e + c * 8, a);
        load8(p + c * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += b[j];
        *reinterpret_cast<bf16x8*>(o + c * 8) = pack8(a);
        if (out32) store8(out32 + static_cast<size_t>(row) * d + c * 8, a);
    }
}

template <typename CT>
__global__ void meshed_mix_kernel(const float* __restrict__ gates, const CT* __restrict__ c, bf16* __restrict__ out,
                                  float* __restrict__ out32, int levels, size_t per_level, float inv_sqrt_levels) {
    pdl_prologue();
    const size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (idx >= per_level) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < levels; ++l) {
        float a[8], v[8];
        load8(gates + l * per_level + idx, a);
        load8(c + l * per_level + idx, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j] / (1.f + __expf(-a[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv_sqrt_levels;
    if (out) store8(out + idx, acc);
    if (out32) store8(out32 + idx, acc);
}

__global__ void aoa_gate_kernel(const float* __restrict__ ig, bf16* __restrict__ out, float* __restrict__ out32, int R,
                                int d) {
    pdl_prologue();
    const size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (idx >= static_cast<size_t>(R) * d) return;
    const size_t row = idx / d, col = idx % d;
    float a[8], b[8];
    load8(ig + row * 2 * d + col, a);
    load8(ig + row * 2 * d + d + col, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a[j] / (1.f + __expf(-b[j]));
    if (out) store8(out + idx, a);
    if (out32) store8(out32 + idx, a);
}

inline void count_launch() { g_cap_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace

extern "C" int cap_add_layernorm(const void* y, int y_dtype, int ldy, const void* residual, int res_dtype, int ldr,
                                 const float* gamma, const float* beta, float eps, const float* pos, int pos_rows,
                                 const uint8_t* zero_rows, void* out, int ldo, float* out_f32, int ldo32, int rows,
                                 int d, cap_stream_t stream) {
    CAP_REQUIRE(y && gamma && beta && (out || out_f32), "cap_add_layernorm: null pointer");
    CAP_REQUIRE(rows > 0 && d > 0 && d % 8 == 0 && d <= LN_MAX_CHUNKS * 256,
                "cap_add_layernorm: d=%d must be a multiple of 8 and <= %d", d, LN_MAX_CHUNKS * 256);
    CAP_REQUIRE(ldy % 8 == 0 && (!out || ldo % 8 == 0) && (!out_f32 || ldo32 % 4 == 0) && (!residual || ldr % 8 == 0),
                "cap_add_layernorm: leading dimensions must be multiples of 8");
    CAP_REQUIRE(!pos || pos_rows > 0, "cap_add_layernorm: pos_rows must be positive");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    bf16* o16 = static_cast<bf16*>(out);
#define CAP_LN(YT, RT)                                                                                              \
    CAP_LAUNCH((add_layernorm_kernel<YT, RT>), blocks, ROWS_PER_BLOCK * 32, 0, s, static_cast<const YT*>(y), ldy,   \
               static_cast<const RT*>(residual), ldr, gamma, beta, eps, pos, pos_rows, zero_rows, o16, ldo, out_f32, \
               ldo32, rows, d)
    const bool y32 = y_dtype == CAP_F32, r32 = res_dtype == CAP_F32;
    if (y32 && r32) CAP_LN(float, float);
    else if (y32) CAP_LN(float, bf16);
    else if (r32) CAP_LN(bf16, float);
    else CAP_LN(bf16, bf16);
#undef CAP_LN
    count_launch();
    return cap_check_launch("add_layernorm_kernel");
}

extern "C" int cap_feature_mask_cast(const void* feats, int feat_dtype, void* out_bf16, uint8_t* mask, int rows,
                                     int d_feature, cap_stream_t stream) {
    CAP_REQUIRE(feats && mask, "cap_feature_mask_cast: null pointer");
    CAP_REQUIRE(rows > 0 && d_feature > 0 && d_feature % 8 == 0, "cap_feature_mask_cast: d_feature %% 8 != 0");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    if (feat_dtype == CAP_F32)
        CAP_LAUNCH((feature_mask_cast_kernel<float>), blocks, ROWS_PER_BLOCK * 32, 0, s, static_cast<const float*>(feats), static_cast<bf16*>(out_bf16), mask, rows, d_feature);
    else
        CAP_LAUNCH((feature_mask_cast_kernel<bf16>), blocks, ROWS_PER_BLOCK * 32, 0, s, static_cast<const bf16*>(feats), static_cast<bf16*>(out_bf16), mask, rows, d_feature);
    count_launch();
    return cap_check_launch("feature_mask_cast_kernel");
}

extern "C" int cap_geometry_bias(const float* boxes, const float* w_g, const float* b_g, float* g, int B, int n,
                                 int H, int d_g, int trig, cap_stream_t stream) {
    CAP_REQUIRE(boxes && w_g && b_g && g, "cap_geometry_bias: null pointer");
    CAP_REQUIRE(B > 0 && n > 0 && H > 0, "cap_geometry_bias: empty problem");
    CAP_REQUIRE(trig ? (d_g % 8 == 0 && d_g > 0) : d_g == 4, "cap_geometry_bias: d_g=%d invalid for trig=%d", d_g,
                trig);
    const int total = B * n * n;
    CAP_LAUNCH((geometry_bias_kernel), (total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), boxes, w_g, b_g, g, B, n, H, d_g, trig);
    count_launch();
    return cap_check_launch("geometry_bias_kernel");
}

extern "C" int cap_region_grid_mask(const float* boxes, uint8_t* mask, int rows, int grid_size, cap_stream_t stream) {
    CAP_REQUIRE(boxes && mask, "cap_region_grid_mask: null pointer");
    CAP_REQUIRE(rows > 0 && grid_size > 0 && grid_size <= 64, "cap_region_grid_mask: bad shape");
    const int total = rows * grid_size * grid_size;
    CAP_LAUNCH((region_grid_mask_kernel), (total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), boxes, mask, rows, grid_size);
    count_launch();
    return cap_check_launch("region_grid_mask_kernel");
}

extern "C" int cap_embed_tokens(const int32_t* tokens, const void* word_emb_bf16, const float* pos_table,
                                int position, int pad_idx, void* out, float* out_f32, uint8_t* padflag_out, int R,
                                int d, cap_stream_t stream) {
    CAP_REQUIRE(tokens && word_emb_bf16 && pos_table && out, "cap_embed_tokens: null pointer");
    CAP_REQUIRE(R > 0 && d % 8 == 0, "cap_embed_tokens: bad shape");
    CAP_LAUNCH((embed_tokens_kernel), (R + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, ROWS_PER_BLOCK * 32, 0, static_cast<cudaStream_t>(stream), tokens, static_cast<const bf16*>(word_emb_bf16), pos_table, position, pad_idx, static_cast<bf16*>(out), out_f32, padflag_out, R, d);
    count_launch();
    return cap_check_launch("embed_tokens_kernel");
}

extern "C" int cap_meshed_mix(const float* gates, const void* c, int c_dtype, void* out, float* out_f32, int levels,
                              int R, int d, cap_stream_t stream) {
    CAP_REQUIRE(gates && c && (out || out_f32) && levels > 0 && R > 0 && d % 8 == 0, "cap_meshed_mix: bad arguments");
    const size_t per_level = static_cast<size_t>(R) * d;
    const int threads = 256;
    const int blocks = static_cast<int>((per_level / 8 + threads - 1) / threads);
    const float inv = 1.f / sqrtf(static_cast<float>(levels));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (c_dtype == CAP_F32)
        CAP_LAUNCH((meshed_mix_kernel<float>), blocks, threads, 0, s, gates, static_cast<const float*>(c),
                   static_cast<bf16*>(out), out_f32, levels, per_level, inv);
    else
        CAP_LAUNCH((meshed_mix_kernel<bf16>), blocks, threads, 0, s, gates, static_cast<const bf16*>(c),
                   static_cast<bf16*>(out), out_f32, levels, per_level, inv);
    count_launch();
    return cap_check_launch("meshed_mix_kernel");
}

extern "C" int cap_aoa_gate(const float* ig, void* out, float* out_f32, int R, int d, cap_stream_t stream) {
    CAP_REQUIRE(ig && (out || out_f32) && R > 0 && d % 8 == 0, "cap_aoa_gate: bad arguments");
    const size_t total = static_cast<size_t>(R) * d;
    const int threads = 256;
    const int blocks = static_cast<int>((total / 8 + threads - 1) / threads);
    CAP_LAUNCH((aoa_gate_kernel), blocks, threads, 0, static_cast<cudaStream_t>(stream), ig, static_cast<bf16*>(out),
               out_f32, R, d);
    count_launch();
    return cap_check_launch("aoa_gate_kernel");
}
