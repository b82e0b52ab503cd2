// CIDEr-D on the host, natively -- plain C++ (compiled by g++), no kernels.
//
// The reward of the self-critical step and the CIDEr of the evaluation loop: reference
// evaluation/cider/cider_scorer.py:9-167 (precook :9-24, document frequencies :80-91, counts2vec :94-119,
// sim :121-148, the per-image average :150-166) driven by evaluation/cider/cider.py:12-38 and
// trainers/vi_trainer.py:137-145.  There every n-gram is a tuple of strings in nested Python dicts; here captions
// arrive as int32 word ids (the Python side numbers the words), an n-gram is 16 bytes, and hypotheses are scored
// in parallel.  Arithmetic is double like the reference's and follows its order of operations (insertion-ordered
// n-grams, per-order norms, clipped dot product, Gaussian length penalty on the BIGRAM count -- counts2vec's
// `if n == 1` tests the zero-based order, reference :115-116), so scores agree to rounding (tests: 1e-12).

#include <cmath>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

#include "cap_common.cuh"

namespace {

constexpr int MAX_ORDER = 4;

struct NGram {
    uint64_t lo, hi;   // (w0+1) | (w1+1) << 32,  (w2+1) | (w3+1) << 32 ; absent positions are 0
    bool operator==(const NGram& o) const { return lo == o.lo && hi == o.hi; }
};

struct NGramHash {
    size_t operator()(const NGram& k) const {
        uint64_t h = k.lo * 0x9E3779B97F4A7C15ull;
        h ^= (k.hi + 0xC2B2AE3D27D4EB4Full) * 0xFF51AFD7ED558CCDull + (h << 6) + (h >> 2);
        h ^= h >> 29;
        return static_cast<size_t>(h * 0xBF58476D1CE4E5B9ull);
    }
};

inline NGram make_ngram(const int32_t* w, int k) {
    NGram g{0, 0};
    g.lo = static_cast<uint64_t>(static_cast<uint32_t>(w[0]) + 1u);
    if (k > 1) g.lo |= static_cast<uint64_t>(static_cast<uint32_t>(w[1]) + 1u) << 32;
    if (k > 2) g.hi = static_cast<uint64_t>(static_cast<uint32_t>(w[2]) + 1u);
    if (k > 3) g.hi |= static_cast<uint64_t>(static_cast<uint32_t>(w[3]) + 1u) << 32;
    return g;
}

using DocFreq = std::unordered_map<NGram, double, NGramHash>;

// precook: the distinct n-grams of a caption in first-occurrence order (orders 1..n, positions left to right),
// with their counts -- the iteration order of the reference's defaultdict.
struct Cooked {
    std::vector<NGram> gram;
    std::vector<int> order;      // zero-based: unigram 0
    std::vector<double> count;
    std::unordered_map<NGram, int, NGramHash> index;
};

void precook(const int32_t* words, int64_t len, int n, Cooked* out) {
    out->gram.clear(); out->order.clear(); out->count.clear(); out->index.clear();
    for (int k = 1; k <= n; ++k)
        for (int64_t i = 0; i + k <= len; ++i) {
            const NGram g = make_ngram(words + i, k);
            auto it = out->index.find(g);
            if (it == out->index.end()) {
                out->index.emplace(g, static_cast<int>(out->gram.size()));
                out->gram.push_back(g);
                out->order.push_back(k - 1);
                out->count.push_back(1.0);
            } else {
                out->count[it->second] += 1.0;
            }
        }
}

// counts2vec: tf-idf weight per n-gram, per-order norms, and the "length" (sum of the bigram counts)
struct Weighted {
    std::vector<double> w;        // aligned with Cooked::gram
    double norm[MAX_ORDER];
    double length;
};

void weigh(const Cooked& c, const DocFreq& df, const double* log_table, int64_t log_table_len, double ref_len, int n,
           Weighted* out) {
    out->w.resize(c.gram.size());
    double sq[MAX_ORDER] = {0, 0, 0, 0};
    out->length = 0;
    volatile double two = 2.0;    // pow(x, 2) through libm as Python's pow does, not folded into x * x
    for (size_t i = 0; i < c.gram.size(); ++i) {
        const auto it = df.find(c.gram[i]);
        const double freq = it == df.end() ? 0.0 : it->second;
        // log(max(1, document frequency)): from the caller's table when it covers the count (numpy's log), else libm's
        double log_df = 0.0;
        if (freq > 1.0) {
            const int64_t f = static_cast<int64_t>(freq);
            log_df = (log_table && f < log_table_len) ? log_table[f] : std::log(freq);
        }
        const double v = c.count[i] * (ref_len - log_df);
        out->w[i] = v;
        sq[c.order[i]] += std::pow(v, two);
        if (c.order[i] == 1) out->length += c.count[i];
    }
    for (int k = 0; k < n; ++k) out->norm[k] = std::sqrt(sq[k]);
    for (int k = n; k < MAX_ORDER; ++k) out->norm[k] = 0;
}

}  // namespace

struct cap_cider {
    int n = 4;
    double sigma = 6.0;
    bool has_corpus = false;
    DocFreq doc_freq;
    double ref_len = 0.0;
    std::vector<double> log_table;   // log_table[c] = log(c) as the caller's numpy computes it
};

extern "C" int cap_cider_create(int n, double sigma, cap_cider** out) {
    CAP_REQUIRE(out != nullptr && n >= 1 && n <= MAX_ORDER && sigma > 0, "cap_cider_create: n must be 1..4 and sigma positive");
    cap_cider* c = new cap_cider;
    c->n = n;
    c->sigma = sigma;
    *out = c;
    return CAP_OK;
}

extern "C" int cap_cider_destroy(cap_cider* c) {
    delete c;
    return CAP_OK;
}

static int check_ragged(const char* who, const int32_t* tokens, const int64_t* caption_offsets, int64_t n_captions) {
    CAP_REQUIRE(caption_offsets != nullptr && caption_offsets[0] == 0, "%s: caption offsets must start at 0", who);
    for (int64_t i = 0; i < n_captions; ++i)
        CAP_REQUIRE(caption_offsets[i + 1] >= caption_offsets[i], "%s: caption offsets decrease at %lld", who, static_cast<long long>(i));
    CAP_REQUIRE(tokens != nullptr || caption_offsets[n_captions] == 0, "%s: null tokens", who);
    for (int64_t i = 0; i < caption_offsets[n_captions]; ++i)
        CAP_REQUIRE(tokens[i] >= 0, "%s: negative word id at %lld", who, static_cast<long long>(i));
    return CAP_OK;
}

// document frequency: in how many images' reference sets an n-gram occurs (reference cider_scorer.py:80-91)
static void count_documents(const int32_t* tokens, const int64_t* caption_offsets, const int64_t* image_offsets,
                            int64_t n_images, int n, DocFreq* df) {
    Cooked cooked;
    std::unordered_map<NGram, int, NGramHash> seen;
    for (int64_t im = 0; im < n_images; ++im) {
        seen.clear();
        for (int64_t c = image_offsets[im]; c < image_offsets[im + 1]; ++c) {
            precook(tokens + caption_offsets[c], caption_offsets[c + 1] - caption_offsets[c], n, &cooked);
            for (const NGram& g : cooked.gram)
                if (seen.emplace(g, 1).second) (*df)[g] += 1.0;
        }
    }
}

extern "C" int cap_cider_set_corpus(cap_cider* c, const int32_t* tokens, const int64_t* caption_offsets,
                                    const int64_t* image_offsets, int64_t n_images, double ref_len,
                                    const double* log_table, int64_t log_table_len) {
    CAP_REQUIRE(c != nullptr && image_offsets != nullptr && n_images > 0, "cap_cider_set_corpus: bad arguments");
    CAP_REQUIRE(image_offsets[0] == 0, "cap_cider_set_corpus: image offsets must start at 0");
    for (int64_t i = 0; i < n_images; ++i)
        CAP_REQUIRE(image_offsets[i + 1] >= image_offsets[i], "cap_cider_set_corpus: image offsets decrease at %lld", static_cast<long long>(i));
    CAP_PROPAGATE(check_ragged("cap_cider_set_corpus", tokens, caption_offsets, image_offsets[n_images]));
    c->doc_freq.clear();
    count_documents(tokens, caption_offsets, image_offsets, n_images, c->n, &c->doc_freq);
    c->ref_len = ref_len;
    c->log_table.assign(log_table, log_table + (log_table ? log_table_len : 0));
    c->has_corpus = true;
    return CAP_OK;
}

extern "C" int64_t cap_cider_max_doc_freq(const cap_cider* c) {
    double best = 0;
    if (c)
        for (const auto& kv : c->doc_freq) best = kv.second > best ? kv.second : best;
    return static_cast<int64_t>(best);
}

extern "C" int cap_cider_score(const cap_cider* c, const int32_t* hyp_tokens, const int64_t* hyp_offsets, int64_t n_hyp,
                               const int32_t* ref_tokens, const int64_t* ref_caption_offsets,
                               const int64_t* ref_group_offsets, double batch_ref_len, const double* log_table,
                               int64_t log_table_len, double* scores, int threads) {
    CAP_REQUIRE(c != nullptr && n_hyp >= 0 && (n_hyp == 0 || scores != nullptr), "cap_cider_score: bad arguments");
    if (n_hyp == 0) return CAP_OK;
    CAP_REQUIRE(ref_group_offsets != nullptr && ref_group_offsets[0] == 0, "cap_cider_score: reference groups must start at 0");
    for (int64_t i = 0; i < n_hyp; ++i)
        CAP_REQUIRE(ref_group_offsets[i + 1] > ref_group_offsets[i], "cap_cider_score: hypothesis %lld has no reference captions",
                    static_cast<long long>(i));
    CAP_PROPAGATE(check_ragged("cap_cider_score (hypotheses)", hyp_tokens, hyp_offsets, n_hyp));
    CAP_PROPAGATE(check_ragged("cap_cider_score (references)", ref_tokens, ref_caption_offsets, ref_group_offsets[n_hyp]));

    // without a corpus the batch's own references are the documents (Cider() built with gts=None, reference cider.py:22-26)
    DocFreq batch_df;
    const DocFreq* df = &c->doc_freq;
    double ref_len = c->ref_len;
    const double* table = c->log_table.empty() ? nullptr : c->log_table.data();
    int64_t table_len = static_cast<int64_t>(c->log_table.size());
    if (!c->has_corpus) {
        count_documents(ref_tokens, ref_caption_offsets, ref_group_offsets, n_hyp, c->n, &batch_df);
        df = &batch_df;
        ref_len = batch_ref_len;
        table = log_table;
        table_len = log_table ? log_table_len : 0;
    }

    const int n = c->n;
    const double sigma = c->sigma;
    auto work = [=](int64_t lo, int64_t hi) {
        Cooked hyp, ref;
        Weighted hw, rw;
        volatile double e = M_E, two = 2.0;
        for (int64_t i = lo; i < hi; ++i) {
            precook(hyp_tokens + hyp_offsets[i], hyp_offsets[i + 1] - hyp_offsets[i], n, &hyp);
            weigh(hyp, *df, table, table_len, ref_len, n, &hw);
            double score[MAX_ORDER] = {0, 0, 0, 0};
            const int64_t r0 = ref_group_offsets[i], r1 = ref_group_offsets[i + 1];
            for (int64_t r = r0; r < r1; ++r) {
                precook(ref_tokens + ref_caption_offsets[r], ref_caption_offsets[r + 1] - ref_caption_offsets[r], n, &ref);
                weigh(ref, *df, table, table_len, ref_len, n, &rw);
                double val[MAX_ORDER] = {0, 0, 0, 0};
                for (size_t g = 0; g < hyp.gram.size(); ++g) {            // clipped dot product, hypothesis order
                    const auto it = ref.index.find(hyp.gram[g]);
                    const double vr = it == ref.index.end() ? 0.0 : rw.w[it->second];
                    const double vh = hw.w[g];
                    val[hyp.order[g]] += (vr < vh ? vr : vh) * vr;
                }
                const double delta = hw.length - rw.length;
                const double penalty = std::pow(e, -std::pow(delta, two) / (2 * std::pow(sigma, two)));
                for (int k = 0; k < n; ++k) {
                    if (hw.norm[k] != 0 && rw.norm[k] != 0) val[k] /= hw.norm[k] * rw.norm[k];
                    score[k] += val[k] * penalty;
                }
            }
            double mean = 0;
            for (int k = 0; k < n; ++k) mean += score[k];
            mean /= n;
            mean /= static_cast<double>(r1 - r0);
            scores[i] = mean * 10.0;
        }
    };
    threads = threads < 1 ? 1 : threads;
    if (threads == 1 || n_hyp < 2 * threads) {
        work(0, n_hyp);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, n_hyp * t / threads, n_hyp * (t + 1) / threads);
        for (auto& th : pool) th.join();
    }
    return CAP_OK;
}
