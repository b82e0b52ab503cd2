// Host glue either side of the caption path -- plain host C++ (compiled by g++), no kernels:
//
//   * cap_host_collate_*: the zero-padding collate of per-image feature rows into one (B, n_max, D) batch,
//     written straight into the (pinned) buffer the H2D copy reads -- replaces InstanceList.__init__ /
//     pad_values (reference utils/instance.py:32-55, 156-171) and the later .to(device) staging copy.
//     The bf16 variant also does the fp32 -> bf16 conversion the engine wants (round to nearest even, the
//     rounding torch's .to(torch.bfloat16) applies), so the host touches every feature exactly once.
//   * cap_vocab_*: caption ids -> text, i.e. Vocab.decode_caption (reference data_utils/vocab.py:104-122)
//     and the consecutive-duplicate collapse the trainers apply to it (reference trainers/vi_trainer.py:251,
//     itertools.groupby over the words).
//
// Once the GPU path runs at ~85 k captions/s these two Python loops are what bounds a caller (1.7 M token
// iterations and 8.5 GB of fp32 features per second), hence native code.

#include <immintrin.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cap_common.cuh"

namespace {

inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;   // NaN -> the canonical quiet NaN (as c10::BFloat16)
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}

// 8 floats per step: same integer recipe as above on AVX2 lanes, non-temporal stores (the batch is written once and
// next read by the DMA engine, so it should not displace the source rows from the cache).  dst 32-byte aligned.
__attribute__((target("avx2"))) inline __m256i eight_to_bf16_avx2(const float* p) {
    const __m256i abs_mask = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000);
    const __m256i bias = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1), qnan = _mm256_set1_epi32(0x7fc0);
    const __m256i u = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i is_nan = _mm256_cmpgt_epi32(_mm256_and_si256(u, abs_mask), inf);
    const __m256i lsb = _mm256_and_si256(_mm256_srli_epi32(u, 16), one);
    const __m256i r = _mm256_srli_epi32(_mm256_add_epi32(u, _mm256_add_epi32(bias, lsb)), 16);
    return _mm256_blendv_epi8(r, qnan, is_nan);
}

__attribute__((target("avx2"))) void row_to_bf16_avx2(const float* src, uint16_t* dst, int n) {
    int c = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    for (; c + 16 <= n; c += 16) {
        const __m256i packed = _mm256_permute4x64_epi64(_mm256_packus_epi32(eight_to_bf16_avx2(src + c), eight_to_bf16_avx2(src + c + 8)), 0xD8);
        if (aligned)
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + c), packed);
        else
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + c), packed);
    }
    for (; c < n; ++c) dst[c] = f32_to_bf16_rne(src[c]);
}

void row_to_bf16_scalar(const float* src, uint16_t* dst, int n) {
    for (int c = 0; c < n; ++c) dst[c] = f32_to_bf16_rne(src[c]);
}

template <typename Fn>
void parallel_over(int items, int threads, Fn fn) {
    threads = std::max(1, std::min(threads, items));
    if (threads == 1) {
        fn(0, items);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int w = 0; w < threads; ++w) {
        const int lo = static_cast<int>(static_cast<int64_t>(items) * w / threads);
        const int hi = static_cast<int>(static_cast<int64_t>(items) * (w + 1) / threads);
        pool.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& t : pool) t.join();
}

int check_collate(const void* rows, const int32_t* n_rows, int B, int n_max, int D, const void* out) {
    CAP_REQUIRE(B >= 0 && n_max >= 0 && D > 0, "cap_host_collate: bad shape B=%d n_max=%d D=%d", B, n_max, D);
    CAP_REQUIRE(B == 0 || (rows != nullptr && n_rows != nullptr && out != nullptr), "cap_host_collate: null pointer");
    for (int i = 0; i < B; ++i)
        CAP_REQUIRE(n_rows[i] >= 0 && n_rows[i] <= n_max, "cap_host_collate: image %d has %d rows, batch holds %d", i,
                    n_rows[i], n_max);
    return CAP_OK;
}

}  // namespace

extern "C" int cap_host_collate_bf16(const float* const* rows, const int32_t* n_rows, int B, int n_max, int D,
                                     uint16_t* out, int threads) {
    CAP_PROPAGATE(check_collate(rows, n_rows, B, n_max, D, out));
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    void (*convert_row)(const float*, uint16_t*, int) = has_avx2 ? row_to_bf16_avx2 : row_to_bf16_scalar;
    // work items are (image, row): images differ in length, rows do not
    parallel_over(B * n_max, threads, [=](int lo, int hi) {
        for (int item = lo; item < hi; ++item) {
            const int i = item / n_max, r = item - i * n_max;
            uint16_t* dst = out + static_cast<size_t>(item) * D;
            if (r >= n_rows[i]) {
                memset(dst, 0, sizeof(uint16_t) * D);
                continue;
            }
            const float* src = rows[i] + static_cast<size_t>(r) * D;
            convert_row(src, dst, D);
        }
        if (has_avx2) _mm_sfence();   // the streaming stores are ordered before the thread is joined
    });
    return CAP_OK;
}

extern "C" int cap_host_collate_f32(const float* const* rows, const int32_t* n_rows, int B, int n_max, int D,
                                    float* out, int threads) {
    CAP_PROPAGATE(check_collate(rows, n_rows, B, n_max, D, out));
    parallel_over(B, threads, [=](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            float* dst = out + static_cast<size_t>(i) * n_max * D;
            const size_t live = static_cast<size_t>(n_rows[i]) * D;
            if (live) memcpy(dst, rows[i], live * sizeof(float));
            memset(dst + live, 0, (static_cast<size_t>(n_max) * D - live) * sizeof(float));
        }
    });
    return CAP_OK;
}

// ---------------------------------------------------------------------------------------------------------
struct cap_vocab {
    std::string blob;               // all words back to back
    std::vector<int64_t> offset;    // V + 1 byte offsets into blob
    std::vector<uint8_t> special;   // itos[i] in specials  (never emitted)
    int64_t eos_idx;
};

extern "C" int cap_vocab_create(const char* words, const int64_t* offsets, int64_t n_words, const uint8_t* is_special,
                                int64_t eos_idx, cap_vocab** out) {
    CAP_REQUIRE(out != nullptr && offsets != nullptr && is_special != nullptr && n_words > 0, "cap_vocab_create: bad arguments");
    CAP_REQUIRE(offsets[0] == 0, "cap_vocab_create: offsets must start at 0");
    for (int64_t i = 0; i < n_words; ++i)
        CAP_REQUIRE(offsets[i + 1] >= offsets[i], "cap_vocab_create: offsets decrease at word %lld", static_cast<long long>(i));
    CAP_REQUIRE(words != nullptr || offsets[n_words] == 0, "cap_vocab_create: null word table");
    cap_vocab* v = new cap_vocab;
    v->blob.assign(words ? words : "", static_cast<size_t>(offsets[n_words]));
    v->offset.assign(offsets, offsets + n_words + 1);
    v->special.assign(is_special, is_special + n_words);
    v->eos_idx = eos_idx;
    *out = v;
    return CAP_OK;
}

extern "C" int cap_vocab_destroy(cap_vocab* v) {
    delete v;
    return CAP_OK;
}

// One caption: the words of ids[0..T) that are not special, up to and including the first eos (which, being
// special, is not emitted); with `collapse`, a word equal to the previously emitted word is dropped (groupby
// over the emitted words).  Returns the byte length; writes when dst != nullptr.
static int64_t decode_one(const cap_vocab* v, const int64_t* ids, int T, bool collapse, char* dst) {
    int64_t len = 0, prev = -1;
    bool first = true;
    const char* blob = v->blob.data();
    for (int t = 0; t < T; ++t) {
        const int64_t id = ids[t];
        if (!v->special[id]) {
            const int64_t lo = v->offset[id], n = v->offset[id + 1] - lo;
            // equal words, not equal ids: two vocabulary entries never spell the same word (stoi is a bijection)
            const bool repeat = collapse && prev >= 0 && (prev == id || (v->offset[prev + 1] - v->offset[prev] == n &&
                                                                        memcmp(blob + v->offset[prev], blob + lo, n) == 0));
            if (!repeat) {
                if (!first) {
                    if (dst) dst[len] = ' ';
                    ++len;
                }
                if (dst) memcpy(dst + len, blob + lo, n);
                len += n;
                first = false;
                prev = id;
            }
        }
        if (id == v->eos_idx) break;
    }
    return len;
}

extern "C" int cap_vocab_decode(const cap_vocab* v, const int64_t* ids, int64_t n_captions, int T, int collapse_repeats,
                                char* out, int64_t out_capacity, int64_t* out_bytes) {
    CAP_REQUIRE(v != nullptr && out_bytes != nullptr && n_captions >= 0 && T >= 0, "cap_vocab_decode: bad arguments");
    CAP_REQUIRE(n_captions == 0 || ids != nullptr, "cap_vocab_decode: null ids");
    const int64_t V = static_cast<int64_t>(v->special.size());
    for (int64_t i = 0; i < n_captions * T; ++i)
        CAP_REQUIRE(ids[i] >= 0 && ids[i] < V, "cap_vocab_decode: id %lld at caption %lld position %lld is outside the vocabulary [0,%lld)",
                    static_cast<long long>(ids[i]), static_cast<long long>(i / T), static_cast<long long>(i % T),
                    static_cast<long long>(V));
    // captions are separated by '\n' (no word holds one: words come out of str.split())
    int64_t need = 0;
    for (int64_t i = 0; i < n_captions; ++i) need += decode_one(v, ids + i * T, T, collapse_repeats != 0, nullptr) + 1;
    *out_bytes = need;
    if (need > out_capacity || (need > 0 && out == nullptr))
        return cap_set_error(CAP_ERR_INVALID, "cap_vocab_decode: output needs %lld bytes, buffer holds %lld",
                             static_cast<long long>(need), static_cast<long long>(out_capacity));
    int64_t at = 0;
    for (int64_t i = 0; i < n_captions; ++i) {
        at += decode_one(v, ids + i * T, T, collapse_repeats != 0, out + at);
        out[at++] = '\n';
    }
    return CAP_OK;
}
