// Whole-path engine: encoder_forward once, then max_len decode steps with the beam state machine
// (models/base_transformer.py:30-53 + models/modules/beam_search.py:85-118), all on one stream,
// no host synchronisation inside, optionally replayed from a CUDA graph.
//
// Algorithmic differences from the reference as written (results are the same up to rounding):
//   * cross-attention K/V are projected ONCE per image per decoder layer (and per encoder level
//     for the meshed decoder) instead of per beam row per step (decoders.py:23,56);
//   * self-attention K/V are projected ONCE per generated token and kept post-projection; the
//     reference caches raw inputs and re-projects all of them each step (attentions.py:297-304);
//   * beam reordering never copies the KV cache: rows read their history through an ancestry table
//     (beam.cu) instead of BeamSearch._expand_state's gathers (beam_search.py:19-34);
//   * step 0 runs all `beam` rows (replicas of <bos>) and masks beams > 0 out of the selection, so
//     every step has the same shape.
#include "cap_common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

// Device weights (bf16 projections, fp32 biases / LayerNorm parameters / tables).  One set serves every engine
// created from the same parent (cap_engine_create_shared): N engines pipelining N batches then read ONE copy of the
// weights -- 48 MB that stay L2-resident under the evict_last hints -- instead of N distinct address ranges.
struct WeightOwner {
    std::vector<void*> allocations;
    cap_fused_weights* stacked = nullptr;   // stacked chain weights (decode_fused.cu), built by the first engine that reserves
    cap_enc_chain_weights* enc_stacked = nullptr;   // the encoder's
    std::mutex lock;
    ~WeightOwner() {
        if (stacked) cap_fused_weights_destroy(stacked);
        if (enc_stacked) cap_enc_chain_weights_destroy(enc_stacked);
        for (void* p : allocations) cudaFree(p);
    }
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    size_t numel() const { return data.size(); }
};

struct Linear {
    bf16* w = nullptr;   // [out, in]
    float* b = nullptr;  // [out] or null
    int out = 0, in = 0;
};

struct Norm {
    float *g = nullptr, *b = nullptr;
};

struct AttentionW {
    Linear qkv;  // stacked fc_q|fc_k|fc_v  [3*hd, d]   (self-attention)
    Linear q;    // fc_q alone               [hd, d]     (cross-attention)
    Linear kv;   // stacked fc_k|fc_v        [2*hd, d]   (cross-attention)
    Linear o;    // fc_o                     [d, hd]
    Norm ln;
    bf16 *mem_k = nullptr, *mem_v = nullptr;  // [n_mem, hd], pre-scaled
    Linear aoa;  // stacked informative|gated [2*d, 2*d]
};

struct FeedForwardW {
    Linear fc1, fc2;
    Norm ln;
};

struct EncoderLayerW {
    AttentionW att;
    FeedForwardW ffn;
};

struct DecoderLayerW {
    AttentionW self_att, cross_att;
    FeedForwardW ffn;
    std::vector<Linear> alphas;  // meshed gates, each [d, 2d]
};

}  // namespace

struct cap_engine {
    cap_model_desc desc;
    std::map<std::string, HostTensor> host;
    std::vector<void*> allocations;            // workspaces, caches (this engine's own)
    std::shared_ptr<WeightOwner> weights;      // device weights, possibly shared with sibling engines
    bool alloc_weights = false;                // dev_alloc target: the weight set (finalize) or the workspaces (reserve)
    bool finalized = false;

    // weights
    Linear vis_proj, vocab_fc;
    Norm enc_ln;
    std::vector<EncoderLayerW> enc;
    std::vector<DecoderLayerW> dec;
    float *geo_w = nullptr, *geo_b = nullptr;
    int d_g = 4;
    bf16* word_emb = nullptr;
    float* word_pos = nullptr;  // [max_len+1, d]

    // reservation
    int max_batch = 0, n_tokens = 0, beam = 0;
    int cur_batch = 0, cur_n = 0;
    bool encoded = false;
    float* vis_pos = nullptr;      // [n_tokens, d]
    void* feat_stage = nullptr;    // H2D landing buffer (fp32-sized)
    float* box_stage = nullptr;
    bf16* feat_bf16 = nullptr;     // [rows_enc, d_feature]
    uint8_t* enc_mask = nullptr;   // [rows_enc]
    float* geometry = nullptr;     // [B, H, n, n]
    bf16* enc_levels = nullptr;    // [enc_layers][rows_enc][d]
    bf16* cross_kv = nullptr;      // [dec_layers][levels][rows_enc][2*hd]
    bf16* qkv_cache = nullptr;     // [dec_layers][T][R][3*hd]
    uint8_t* padflag = nullptr;    // [T][R]
    // scratch shared by encoder and decoder phases
    bf16 *buf_x = nullptr, *buf_a = nullptr, *buf_att = nullptr, *buf_h = nullptr, *buf_q = nullptr, *buf_cat = nullptr;
    bf16* buf_qkv = nullptr;       // encoder [rows_enc][3*hd]
    bf16* buf_c = nullptr;         // decoder [levels][R][d]
    bf16* buf_mix = nullptr;       // decoder [R][d] meshed mix
    // fp32 twins of the residual stream (x -> a -> c -> x): only GEMM operands are rounded to bf16
    float *res_x = nullptr, *res_a = nullptr, *res_c = nullptr, *res_mix = nullptr;
    float* buf_y32 = nullptr;
    float* logits = nullptr;
    int ld_logits = 0;
    float* part_ms = nullptr;  // vocabulary GEMM chunk statistics [R][chunks][2]
    int vocab_chunks = 0;
    int64_t* out_ids = nullptr;
    float* out_logp = nullptr;
    cap_beam* beam_state = nullptr;
    bool fuse_ln = true;                 // Linear + residual + LayerNorm as one cluster kernel
    cap_fused_decoder* fused = nullptr;  // GEMM chains of the decode step (decode_fused.cu) when the model is covered
    cap_enc_chains* enc_chains = nullptr;  // GEMM chains of the encoder, ditto

    // CUDA graph of a full beam search (begin + T steps + finalize)
    cudaGraphExec_t graph_exec = nullptr;
    cudaStream_t capture_stream = nullptr;  // the legacy default stream cannot be captured: capture here, replay anywhere
    int graph_batch = 0, graph_n = 0, graph_out_size = 0;   // the captured launches bake B, n and out_size into their arguments
    int64_t* graph_ids = nullptr;
    float* graph_logp = nullptr;
    bool warmed = false;

    int hd() const { return desc.heads * desc.d_k; }
    int levels() const { return desc.decoder_kind == CAP_DEC_MESHED ? desc.n_enc_levels : 1; }
};

namespace {

template <typename T>
int dev_alloc(cap_engine* e, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, count * sizeof(T) + 16);
    if (err != cudaSuccess)
        return cap_set_error(CAP_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T),
                             cudaGetErrorString(err));
    (e->alloc_weights ? e->weights->allocations : e->allocations).push_back(p);
    *out = static_cast<T*>(p);
    return CAP_OK;
}

const HostTensor* find(cap_engine* e, const std::string& name) {
    auto it = e->host.find(name);
    return it == e->host.end() ? nullptr : &it->second;
}

int upload_f32(cap_engine* e, const std::vector<float>& v, float** out) {
    CAP_PROPAGATE(dev_alloc(e, out, v.size()));
    CAP_CHECK_CUDA(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return CAP_OK;
}

int upload_bf16(cap_engine* e, const std::vector<float>& v, bf16** out) {
    std::vector<bf16> tmp(v.size());
    for (size_t i = 0; i < v.size(); ++i) tmp[i] = __float2bfloat16_rn(v[i]);
    CAP_PROPAGATE(dev_alloc(e, out, v.size()));
    CAP_CHECK_CUDA(cudaMemcpy(*out, tmp.data(), v.size() * 2, cudaMemcpyHostToDevice));
    return CAP_OK;
}

// Stack several nn.Linear layers (same `in`) along the output dimension.
int make_linear(cap_engine* e, const std::vector<std::string>& prefixes, int in, bool need_bias, Linear* out) {
    std::vector<float> w, b;
    int total_out = 0;
    bool any_bias = false;
    for (const std::string& p : prefixes) {
        const HostTensor* wt = find(e, p + ".weight");
        if (!wt) return cap_set_error(CAP_ERR_STATE, "missing weight '%s.weight'", p.c_str());
        if (wt->shape.size() != 2 || wt->shape[1] != in)
            return cap_set_error(CAP_ERR_INVALID, "'%s.weight' has the wrong shape (expected [*, %d])", p.c_str(), in);
        const int o = static_cast<int>(wt->shape[0]);
        w.insert(w.end(), wt->data.begin(), wt->data.end());
        const HostTensor* bt = find(e, p + ".bias");
        if (bt) {
            if (static_cast<int>(bt->numel()) != o)
                return cap_set_error(CAP_ERR_INVALID, "'%s.bias' has the wrong length", p.c_str());
            b.insert(b.end(), bt->data.begin(), bt->data.end());
            any_bias = true;
        } else {
            if (need_bias) return cap_set_error(CAP_ERR_STATE, "missing weight '%s.bias'", p.c_str());
            b.insert(b.end(), o, 0.f);
        }
        total_out += o;
    }
    out->out = total_out;
    out->in = in;
    CAP_PROPAGATE(upload_bf16(e, w, &out->w));
    if (any_bias) CAP_PROPAGATE(upload_f32(e, b, &out->b));
    return CAP_OK;
}

int make_norm(cap_engine* e, const std::string& prefix, int d, Norm* out) {
    const HostTensor* g = find(e, prefix + ".weight");
    const HostTensor* b = find(e, prefix + ".bias");
    if (!g || !b) return cap_set_error(CAP_ERR_STATE, "missing LayerNorm '%s'", prefix.c_str());
    if (static_cast<int>(g->numel()) != d || static_cast<int>(b->numel()) != d)
        return cap_set_error(CAP_ERR_INVALID, "LayerNorm '%s' has the wrong size", prefix.c_str());
    CAP_PROPAGATE(upload_f32(e, g->data, &out->g));
    CAP_PROPAGATE(upload_f32(e, b->data, &out->b));
    return CAP_OK;
}

int make_attention(cap_engine* e, const std::string& p, bool self_att, int att_kind, bool aoa, AttentionW* out) {
    const cap_model_desc& m = e->desc;
    const int d = m.d_model, hd = e->hd();
    const std::string a = p + ".attention";
    if (self_att) {
        CAP_PROPAGATE(make_linear(e, {a + ".fc_q", a + ".fc_k", a + ".fc_v"}, d, true, &out->qkv));
    } else {
        CAP_PROPAGATE(make_linear(e, {a + ".fc_q"}, d, true, &out->q));
        CAP_PROPAGATE(make_linear(e, {a + ".fc_k", a + ".fc_v"}, d, true, &out->kv));
    }
    CAP_PROPAGATE(make_linear(e, {a + ".fc_o"}, hd, true, &out->o));
    CAP_PROPAGATE(make_norm(e, p + ".layer_norm", d, &out->ln));
    if (att_kind == CAP_ATT_MEMORY) {
        const HostTensor* mk = find(e, a + ".m_k");
        const HostTensor* mv = find(e, a + ".m_v");
        if (!mk || !mv) return cap_set_error(CAP_ERR_STATE, "missing memory slots under '%s'", a.c_str());
        const size_t want = static_cast<size_t>(m.n_memory) * hd;
        if (mk->numel() != want || mv->numel() != want)
            return cap_set_error(CAP_ERR_INVALID, "memory slots under '%s' have the wrong size", a.c_str());
        std::vector<float> sk(mk->data), sv(mv->data);
        const float fk = std::sqrt(static_cast<float>(m.d_k)), fv = std::sqrt(static_cast<float>(m.n_memory));
        for (float& x : sk) x *= fk;  // attentions.py:171
        for (float& x : sv) x *= fv;  // attentions.py:172
        CAP_PROPAGATE(upload_bf16(e, sk, &out->mem_k));
        CAP_PROPAGATE(upload_bf16(e, sv, &out->mem_v));
    }
    if (aoa) CAP_PROPAGATE(make_linear(e, {p + ".informative_attention", p + ".gated_attention"}, 2 * d, true, &out->aoa));
    return CAP_OK;
}

int make_ffn(cap_engine* e, const std::string& p, FeedForwardW* out) {
    const cap_model_desc& m = e->desc;
    CAP_PROPAGATE(make_linear(e, {p + ".fc1"}, m.d_model, true, &out->fc1));
    CAP_PROPAGATE(make_linear(e, {p + ".fc2"}, m.d_ff, true, &out->fc2));
    CAP_PROPAGATE(make_norm(e, p + ".layer_norm", m.d_model, &out->ln));
    return CAP_OK;
}

int run_linear(const bf16* x, int ldx, const Linear& l, void* y, int ldy, int out_dtype, int act, int M,
               cudaStream_t s) {
    return cap_linear(x, ldx, l.w, l.b, y, ldy, out_dtype, act, M, l.out, l.in, s);
}

// LN(res32 + y32): bf16 copy for the next GEMM, fp32 copy for the next residual (both dense, ld = d)
int run_ln(const float* y32, const float* res32, const Norm& n, const float* pos, int pos_rows,
           const uint8_t* zero_rows, bf16* out16, float* out32, int rows, int d, cudaStream_t s) {
    return cap_add_layernorm(y32, CAP_F32, d, res32, CAP_F32, d, n.g, n.b, 1e-5f, pos, pos_rows, zero_rows, out16, d,
                             out32, d, rows, d, s);
}

// timing-only ablations (results are wrong): OPENVIIC_DBG_ABLATE bit 0 = no decode self-attention kernels,
// bit 1 = no encoder layers, bit 2 = no decode chains, bit 3 = no decode cross-attention kernels
int dbg_ablate() {
    static const int v = getenv("OPENVIIC_DBG_ABLATE") ? atoi(getenv("OPENVIIC_DBG_ABLATE")) : 0;
    return v;
}

// LN(res32 + x.W^T + b): one cluster kernel when the row fits a cluster (d_model 128..1024), else GEMM + LayerNorm
int run_linear_ln(cap_engine* e, const bf16* x, int ldx, const Linear& l, const float* res32, const Norm& n,
                  const float* pos, int pos_rows, const uint8_t* zero_rows, bf16* out16, float* out32, int rows,
                  cudaStream_t s) {
    const int d = l.out;
    if (e->fuse_ln && d % 128 == 0 && d <= 1024)
        return cap_linear_layernorm(x, ldx, l.w, l.b, res32, d, n.g, n.b, 1e-5f, pos, pos_rows, zero_rows, out16, d, out32,
                                    d, rows, d, l.in, s);
    CAP_PROPAGATE(run_linear(x, ldx, l, e->buf_y32, d, CAP_F32, CAP_ACT_NONE, rows, s));
    return run_ln(e->buf_y32, res32, n, pos, pos_rows, zero_rows, out16, out32, rows, d, s);
}

// AoA: out = Linear_i([q, a]) * sigmoid(Linear_g([q, a]))   (attentions.py:311-315)
int run_aoa(cap_engine* e, const AttentionW& w, const bf16* queries, const bf16* att_out, bf16* out, float* out32,
            int rows, cudaStream_t s) {
    const int d = e->desc.d_model;
    CAP_CHECK_CUDA(cudaMemcpy2DAsync(e->buf_cat, 2 * d * 2, queries, d * 2, d * 2, rows, cudaMemcpyDeviceToDevice, s));
    CAP_CHECK_CUDA(cudaMemcpy2DAsync(e->buf_cat + d, 2 * d * 2, att_out, d * 2, d * 2, rows, cudaMemcpyDeviceToDevice, s));
    CAP_PROPAGATE(run_linear(e->buf_cat, 2 * d, w.aoa, e->buf_y32, 2 * d, CAP_F32, CAP_ACT_NONE, rows, s));
    return cap_aoa_gate(e->buf_y32, out, out32, rows, d, s);
}

}  // namespace

extern "C" int cap_engine_create(const cap_model_desc* desc, cap_engine** out) {
    CAP_REQUIRE(desc && out, "cap_engine_create: null pointer");
    const cap_model_desc& m = *desc;
    CAP_REQUIRE(m.d_k == 64 && m.d_v == 64, "cap_engine_create: kernels are specialised for d_k = d_v = 64 (got %d/%d)",
                m.d_k, m.d_v);
    CAP_REQUIRE(m.d_model > 0 && m.d_model % 64 == 0 && m.d_model <= 2048, "cap_engine_create: d_model %% 64 != 0");
    CAP_REQUIRE(m.heads > 0 && m.d_ff % 8 == 0 && m.d_feature % 8 == 0, "cap_engine_create: bad head/ff/feature size");
    CAP_REQUIRE(m.enc_layers > 0 && m.dec_layers > 0, "cap_engine_create: need at least one layer");
    CAP_REQUIRE(m.max_len > 0 && m.max_len <= 64, "cap_engine_create: max_len must be in [1,64]");
    CAP_REQUIRE(m.vocab > 8, "cap_engine_create: vocab too small");
    CAP_REQUIRE(m.encoder_kind >= CAP_ENC_PLAIN && m.encoder_kind <= CAP_ENC_GEOMETRIC, "bad encoder_kind");
    CAP_REQUIRE(m.enc_attention >= CAP_ATT_SDPA && m.enc_attention <= CAP_ATT_MEMORY, "bad enc_attention");
    CAP_REQUIRE((m.encoder_kind == CAP_ENC_GEOMETRIC) == (m.enc_attention == CAP_ATT_GEOMETRY),
                "geometric encoder and geometry attention must be selected together");
    if (m.decoder_kind == CAP_DEC_MESHED) {
        CAP_REQUIRE(m.encoder_kind == CAP_ENC_MULTILEVEL, "meshed decoder needs the multi-level encoder");
        CAP_REQUIRE(m.n_enc_levels == m.enc_layers, "meshed decoder: n_enc_levels must equal encoder layers");
    }
    cap_engine* e = new cap_engine();
    e->desc = m;
    e->weights = std::make_shared<WeightOwner>();
    *out = e;
    return CAP_OK;
}

// A second engine over the SAME device weights (no upload, no copy): its own workspaces, caches, beam state and CUDA
// graph, so that it can caption another batch concurrently on another stream.  The weight set lives until the last
// engine that references it is destroyed; `parent` itself may be destroyed first.
extern "C" int cap_engine_create_shared(cap_engine* parent, cap_engine** out) {
    CAP_REQUIRE(parent && out, "cap_engine_create_shared: null pointer");
    CAP_REQUIRE(parent->finalized, "cap_engine_create_shared: finalize the parent first");
    cap_engine* e = new cap_engine();
    e->desc = parent->desc;
    e->weights = parent->weights;
    e->fuse_ln = parent->fuse_ln;
    e->vis_proj = parent->vis_proj; e->vocab_fc = parent->vocab_fc; e->enc_ln = parent->enc_ln;
    e->enc = parent->enc; e->dec = parent->dec;
    e->geo_w = parent->geo_w; e->geo_b = parent->geo_b; e->d_g = parent->d_g;
    e->word_emb = parent->word_emb; e->word_pos = parent->word_pos;
    e->finalized = true;
    *out = e;
    return CAP_OK;
}

extern "C" int cap_engine_destroy(cap_engine* e) {
    if (!e) return CAP_OK;
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->capture_stream) cudaStreamDestroy(e->capture_stream);
    if (e->beam_state) cap_beam_destroy(e->beam_state);
    if (e->fused) cap_fused_destroy(e->fused);
    if (e->enc_chains) cap_enc_chains_destroy(e->enc_chains);
    for (void* p : e->allocations) cudaFree(p);
    delete e;
    return CAP_OK;
}

extern "C" int cap_engine_load_weight(cap_engine* e, const char* name, const float* data_host, const int64_t* shape,
                                      int ndim) {
    CAP_REQUIRE(e && name && data_host && (shape || ndim == 0), "cap_engine_load_weight: null pointer");
    CAP_REQUIRE(!e->finalized, "cap_engine_load_weight: engine already finalized");
    HostTensor t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        n *= static_cast<size_t>(shape[i]);
    }
    t.data.assign(data_host, data_host + n);
    e->host[name] = std::move(t);
    return CAP_OK;
}

extern "C" int cap_engine_finalize(cap_engine* e) {
    CAP_REQUIRE(e != nullptr, "cap_engine_finalize: null engine");
    CAP_REQUIRE(!e->finalized, "cap_engine_finalize: called twice");
    const cap_model_desc& m = e->desc;
    const int d = m.d_model;
    e->alloc_weights = true;
    struct Reset { cap_engine* e; ~Reset() { e->alloc_weights = false; } } reset{e};
    CAP_PROPAGATE(make_linear(e, {"vision_embedding.proj"}, m.d_feature, true, &e->vis_proj));
    CAP_PROPAGATE(make_norm(e, "encoder.layer_norm", d, &e->enc_ln));
    e->enc.resize(m.enc_layers);
    for (int i = 0; i < m.enc_layers; ++i) {
        const std::string p = "encoder.layers." + std::to_string(i);
        CAP_PROPAGATE(make_attention(e, p + ".mhatt", true, m.enc_attention, m.aoa_enc != 0, &e->enc[i].att));
        CAP_PROPAGATE(make_ffn(e, p + ".pwff", &e->enc[i].ffn));
    }
    if (m.encoder_kind == CAP_ENC_GEOMETRIC) {
        e->d_g = m.trig_geometry ? d / m.heads : 4;
        std::vector<float> w, b;
        for (int h = 0; h < m.heads; ++h) {
            const HostTensor* wt = find(e, "encoder.fc_gs." + std::to_string(h) + ".weight");
            const HostTensor* bt = find(e, "encoder.fc_gs." + std::to_string(h) + ".bias");
            if (!wt || !bt) return cap_set_error(CAP_ERR_STATE, "missing encoder.fc_gs.%d", h);
            if (static_cast<int>(wt->numel()) != e->d_g) return cap_set_error(CAP_ERR_INVALID, "encoder.fc_gs.%d wrong size", h);
            w.insert(w.end(), wt->data.begin(), wt->data.end());
            b.push_back(bt->data[0]);
        }
        CAP_PROPAGATE(upload_f32(e, w, &e->geo_w));
        CAP_PROPAGATE(upload_f32(e, b, &e->geo_b));
    }
    e->dec.resize(m.dec_layers);
    for (int i = 0; i < m.dec_layers; ++i) {
        const std::string p = "decoder.layers." + std::to_string(i);
        DecoderLayerW& L = e->dec[i];
        CAP_PROPAGATE(make_attention(e, p + ".self_attn", true, CAP_ATT_SDPA, m.aoa_dec_self != 0, &L.self_att));
        CAP_PROPAGATE(make_attention(e, p + ".enc_attn", false, CAP_ATT_SDPA, m.aoa_dec_cross != 0, &L.cross_att));
        CAP_PROPAGATE(make_ffn(e, p + ".pwff", &L.ffn));
        if (m.decoder_kind == CAP_DEC_MESHED) {
            L.alphas.resize(m.n_enc_levels);
            for (int l = 0; l < m.n_enc_levels; ++l)
                CAP_PROPAGATE(make_linear(e, {p + ".fc_alphas." + std::to_string(l)}, 2 * d, true, &L.alphas[l]));
        }
    }
    {
        const HostTensor* emb = find(e, "decoder.word_emb.components.weight");
        if (!emb) return cap_set_error(CAP_ERR_STATE, "missing decoder.word_emb.components.weight");
        if (emb->shape.size() != 2 || emb->shape[0] != m.vocab || emb->shape[1] != d)
            return cap_set_error(CAP_ERR_INVALID, "decoder.word_emb.components.weight has the wrong shape");
        CAP_PROPAGATE(upload_bf16(e, emb->data, &e->word_emb));
        const HostTensor* pos = find(e, "decoder.pos_emb.weight");
        if (!pos) return cap_set_error(CAP_ERR_STATE, "missing decoder.pos_emb.weight");
        if (pos->shape.size() != 2 || pos->shape[0] < m.max_len + 1 || pos->shape[1] != d)
            return cap_set_error(CAP_ERR_INVALID, "decoder.pos_emb.weight has the wrong shape");
        CAP_PROPAGATE(upload_f32(e, pos->data, &e->word_pos));
    }
    CAP_PROPAGATE(make_linear(e, {"decoder.fc"}, d, false, &e->vocab_fc));
    CAP_REQUIRE(e->vocab_fc.out == m.vocab, "decoder.fc.weight rows (%d) != vocab (%d)", e->vocab_fc.out, m.vocab);
    e->host.clear();
    e->finalized = true;
    return CAP_OK;
}

extern "C" int cap_engine_reserve(cap_engine* e, int max_batch, int n_tokens, int beam) {
    CAP_REQUIRE(e && e->finalized, "cap_engine_reserve: finalize the engine first");
    CAP_REQUIRE(e->max_batch == 0, "cap_engine_reserve: already reserved (create a new engine to resize)");
    CAP_REQUIRE(max_batch > 0 && n_tokens > 0 && beam > 0 && beam <= 8, "cap_engine_reserve: bad sizes");
    const cap_model_desc& m = e->desc;
    const int n_mem = m.enc_attention == CAP_ATT_MEMORY ? m.n_memory : 0;
    CAP_REQUIRE(n_tokens + n_mem <= 160, "cap_engine_reserve: n_tokens + memory slots must be <= 160");
    const int d = m.d_model, hd = e->hd(), T = m.max_len, lv = e->levels();
    const size_t rows_enc = static_cast<size_t>(max_batch) * n_tokens;
    const size_t R = static_cast<size_t>(max_batch) * beam;
    const size_t rows_max = std::max(rows_enc, R * lv);
    e->max_batch = max_batch;
    e->n_tokens = n_tokens;
    e->beam = beam;

    // DETR-style 1-D sinusoid over token index 1..n (models/modules/pos_embeddings.py:58-72)
    {
        std::vector<float> tab(static_cast<size_t>(n_tokens) * d);
        for (int p = 0; p < n_tokens; ++p)
            for (int j = 0; j < d; ++j) {
                const float dim_t = std::pow(10000.0f, 2.0f * static_cast<float>(j / 2) / static_cast<float>(d));
                const float a = static_cast<float>(p + 1) / dim_t;
                tab[static_cast<size_t>(p) * d + j] = (j % 2 == 0) ? std::sin(a) : std::cos(a);
            }
        CAP_PROPAGATE(upload_f32(e, tab, &e->vis_pos));
    }
    float* stage = nullptr;
    CAP_PROPAGATE(dev_alloc(e, &stage, rows_enc * m.d_feature));
    e->feat_stage = stage;
    CAP_PROPAGATE(dev_alloc(e, &e->box_stage, rows_enc * 4));
    CAP_PROPAGATE(dev_alloc(e, &e->feat_bf16, rows_enc * m.d_feature));
    CAP_PROPAGATE(dev_alloc(e, &e->enc_mask, rows_enc));
    if (m.encoder_kind == CAP_ENC_GEOMETRIC)
        CAP_PROPAGATE(dev_alloc(e, &e->geometry, static_cast<size_t>(max_batch) * m.heads * n_tokens * n_tokens));
    CAP_PROPAGATE(dev_alloc(e, &e->enc_levels, static_cast<size_t>(m.enc_layers) * rows_enc * d));
    CAP_PROPAGATE(dev_alloc(e, &e->cross_kv, static_cast<size_t>(m.dec_layers) * lv * rows_enc * 2 * hd));
    CAP_PROPAGATE(dev_alloc(e, &e->qkv_cache, static_cast<size_t>(m.dec_layers) * T * R * 3 * hd));
    CAP_PROPAGATE(dev_alloc(e, &e->padflag, static_cast<size_t>(T) * R));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_x, rows_max * d));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_a, rows_max * d));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_att, rows_max * hd));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_h, rows_max * m.d_ff));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_q, rows_max * hd));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_cat, rows_max * 2 * d));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_qkv, rows_enc * 3 * hd));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_c, R * lv * d));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_mix, R * d));
    CAP_PROPAGATE(dev_alloc(e, &e->res_x, rows_max * d));
    CAP_PROPAGATE(dev_alloc(e, &e->res_a, rows_max * d));
    CAP_PROPAGATE(dev_alloc(e, &e->res_c, R * lv * d));
    CAP_PROPAGATE(dev_alloc(e, &e->res_mix, R * d));
    CAP_PROPAGATE(dev_alloc(e, &e->buf_y32, rows_max * static_cast<size_t>(std::max(2 * d, hd))));
    e->ld_logits = (m.vocab + 31) / 32 * 32;
    CAP_PROPAGATE(dev_alloc(e, &e->logits, R * e->ld_logits));
    e->vocab_chunks = ((m.vocab + 255) / 256) * 8;
    CAP_PROPAGATE(dev_alloc(e, &e->part_ms, R * e->vocab_chunks * 2));
    CAP_PROPAGATE(dev_alloc(e, &e->out_ids, R * T));
    CAP_PROPAGATE(dev_alloc(e, &e->out_logp, R * T));
    CAP_PROPAGATE(cap_beam_create(max_batch, beam, T, m.vocab, m.eos_idx, &e->beam_state));

    // Standard and meshed decoders at the reference's sizes: the decode step runs as GEMM chains (decode_fused.cu) with
    // the stand-alone attention kernels between them.  OPENVIIC_FUSED_DECODE=0: one kernel per operator (the path every
    // other configuration -- attention-on-attention, other widths -- takes).
    const char* env = getenv("OPENVIIC_FUSED_DECODE");
    const bool want_fused = !(env && atoi(env) == 0);
    const bool meshed = m.decoder_kind == CAP_DEC_MESHED;
    if (want_fused && (!meshed || lv == 3) && !m.aoa_dec_self && !m.aoa_dec_cross && d == 512 && m.heads == 8 &&
        m.d_k == 64 && m.d_ff == 2048 && m.dec_layers <= 6 && beam <= 5 && T <= 40 && n_tokens <= 104 &&
        e->vocab_fc.b == nullptr && e->vocab_chunks <= 512) {
        std::vector<cap_fused_layer> layers(m.dec_layers);
        for (int l = 0; l < m.dec_layers; ++l) {
            const DecoderLayerW& L = e->dec[l];
            cap_fused_layer& f = layers[l];
            memset(&f, 0, sizeof(f));
            f.w_qkv = L.self_att.qkv.w; f.b_qkv = L.self_att.qkv.b;
            f.w_o1 = L.self_att.o.w; f.b_o1 = L.self_att.o.b; f.ln1_g = L.self_att.ln.g; f.ln1_b = L.self_att.ln.b;
            f.w_q = L.cross_att.q.w; f.b_q = L.cross_att.q.b;
            f.w_o2 = L.cross_att.o.w; f.b_o2 = L.cross_att.o.b; f.ln2_g = L.cross_att.ln.g; f.ln2_b = L.cross_att.ln.b;
            f.w_fc1 = L.ffn.fc1.w; f.b_fc1 = L.ffn.fc1.b; f.w_fc2 = L.ffn.fc2.w; f.b_fc2 = L.ffn.fc2.b;
            f.ln3_g = L.ffn.ln.g; f.ln3_b = L.ffn.ln.b;
            f.n_levels = meshed ? lv : 0;
            for (int i = 0; meshed && i < lv; ++i) {
                f.w_alpha[i] = L.alphas[i].w;
                f.b_alpha[i] = L.alphas[i].b;
            }
        }
        cap_fused_desc fd = {};
        fd.d_model = d; fd.heads = m.heads; fd.d_ff = m.d_ff; fd.n_layers = m.dec_layers; fd.vocab = m.vocab;
        fd.max_len = T; fd.beam = beam; fd.pad_idx = m.pad_idx; fd.max_rows = static_cast<int>(R);
        fd.layers = layers.data();
        fd.w_vocab = e->vocab_fc.w; fd.word_emb = e->word_emb; fd.word_pos = e->word_pos;
        fd.tokens = cap_beam_tokens(e->beam_state);
        fd.padflag = e->padflag; fd.qkv_cache = e->qkv_cache;
        fd.logits = e->logits; fd.ld_logits = e->ld_logits; fd.part_ms = e->part_ms;
        fd.att_in = e->buf_att; fd.q_out = e->buf_q;   // buf_att: [levels][max_rows][hd]
        {   // one stacked copy of the chain weights per weight set, shared by every engine over it
            std::lock_guard<std::mutex> guard(e->weights->lock);
            if (!e->weights->stacked) CAP_PROPAGATE(cap_fused_weights_create(layers.data(), m.dec_layers, &e->weights->stacked));
            fd.stacked = e->weights->stacked;
        }
        CAP_PROPAGATE(cap_fused_create(&fd, &e->fused));
    }
    // Encoder as chains (decode_fused.cu): vision projection + LayerNorm + positions, then per layer fc_o + LN, FFN + LN
    // and the next q|k|v / the decoder's cross K|V projections, the self-attention kernel between the stages.
    // OPENVIIC_ENC_CHAINS=0: one kernel per operator (also the path of the attention-on-attention encoder).
    {
        const char* enc_env = getenv("OPENVIIC_ENC_CHAINS");
        const bool want = !(enc_env && atoi(enc_env) == 0);
        const int n_kv = m.dec_layers * lv;
        if (want && !m.aoa_enc && d == 512 && hd == 512 && m.d_ff == 2048 && m.enc_layers <= 6 && m.d_feature % 64 == 0 &&
            n_kv <= CAP_ENC_MAX_KV && rows_enc <= (1u << 30)) {
            std::vector<cap_enc_layer> layers(m.enc_layers);
            for (int l = 0; l < m.enc_layers; ++l) {
                const EncoderLayerW& L = e->enc[l];
                cap_enc_layer& f = layers[l];
                f.w_qkv = L.att.qkv.w; f.b_qkv = L.att.qkv.b; f.w_o = L.att.o.w; f.b_o = L.att.o.b;
                f.ln1_g = L.att.ln.g; f.ln1_b = L.att.ln.b;
                f.w_fc1 = L.ffn.fc1.w; f.b_fc1 = L.ffn.fc1.b; f.w_fc2 = L.ffn.fc2.w; f.b_fc2 = L.ffn.fc2.b;
                f.ln2_g = L.ffn.ln.g; f.ln2_b = L.ffn.ln.b;
            }
            cap_enc_chain_desc ed = {};
            ed.d_model = d; ed.d_ff = m.d_ff; ed.d_feature = m.d_feature; ed.n_layers = m.enc_layers;
            ed.max_rows = static_cast<int>(rows_enc);
            ed.layers = layers.data();
            ed.w_vis = e->vis_proj.w; ed.b_vis = e->vis_proj.b; ed.ln0_g = e->enc_ln.g; ed.ln0_b = e->enc_ln.b;
            ed.pos = e->vis_pos;
            ed.n_kv = n_kv;
            for (int l = 0; l < m.dec_layers; ++l)
                for (int i = 0; i < lv; ++i) {
                    const int k = l * lv + i;
                    ed.w_kv[k] = e->dec[l].cross_att.kv.w;
                    ed.b_kv[k] = e->dec[l].cross_att.kv.b;
                    ed.kv_level[k] = (m.decoder_kind == CAP_DEC_MESHED) ? i : m.enc_layers - 1;
                    ed.kv_dst[k] = e->cross_kv + static_cast<size_t>(k) * rows_enc * 2 * hd;
                }
            ed.feats = e->feat_bf16; ed.qkv_out = e->buf_qkv; ed.att_in = e->buf_att;
            ed.levels_out = e->enc_levels; ed.level_stride = rows_enc * d;
            ed.row_mask = e->enc_mask;
            {
                std::lock_guard<std::mutex> guard(e->weights->lock);
                if (!e->weights->enc_stacked) CAP_PROPAGATE(cap_enc_chain_weights_create(&ed, &e->weights->enc_stacked));
                ed.stacked = e->weights->enc_stacked;
            }
            CAP_PROPAGATE(cap_enc_chains_create(&ed, &e->enc_chains));
        }
    }
    return CAP_OK;
}

extern "C" int cap_engine_encode(cap_engine* e, const void* feats, int feat_dtype, const float* boxes, int B, int n,
                                 cap_stream_t stream) {
    CAP_REQUIRE(e && e->max_batch > 0, "cap_engine_encode: reserve the engine first");
    CAP_REQUIRE(feats != nullptr, "cap_engine_encode: null features");
    CAP_REQUIRE(B > 0 && B <= e->max_batch, "cap_engine_encode: batch %d outside (0,%d]", B, e->max_batch);
    CAP_REQUIRE(n > 0 && n <= e->n_tokens, "cap_engine_encode: n=%d outside (0,%d]", n, e->n_tokens);
    const cap_model_desc& m = e->desc;
    CAP_REQUIRE(m.encoder_kind != CAP_ENC_GEOMETRIC || boxes != nullptr, "cap_engine_encode: boxes required");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int d = m.d_model, hd = e->hd(), lv = e->levels();
    const int rows = B * n;
    const size_t rows_cap = static_cast<size_t>(e->max_batch) * e->n_tokens;
    e->cur_batch = B;
    e->cur_n = n;

    // V1: padding mask from the RAW features + cast
    CAP_PROPAGATE(cap_feature_mask_cast(feats, feat_dtype, e->feat_bf16, e->enc_mask, rows, m.d_feature, s));
    if (m.encoder_kind == CAP_ENC_GEOMETRIC)
        CAP_PROPAGATE(cap_geometry_bias(boxes, e->geo_w, e->geo_b, e->geometry, B, n, m.heads, e->d_g,
                                        m.trig_geometry, s));
    auto self_attention = [&](const AttentionW& att) -> int {   // A1-A3 on the fused q|k|v rows of the current layer
        cap_attention_args a = {};
        a.q = e->buf_qkv;
        a.k = e->buf_qkv + hd;
        a.v = e->buf_qkv + 2 * hd;
        a.out = e->buf_att;
        a.q_bs = a.k_bs = a.v_bs = static_cast<int64_t>(n) * 3 * hd;
        a.o_bs = static_cast<int64_t>(n) * hd;
        a.ldq = a.ldk = a.ldv = 3 * hd;
        a.ldo = hd;
        a.mask = e->enc_mask;
        a.mask_bs = n;
        a.mask_qs = 0;
        a.geometry = m.enc_attention == CAP_ATT_GEOMETRY ? e->geometry : nullptr;
        if (m.enc_attention == CAP_ATT_MEMORY) {
            a.mem_k = att.mem_k;
            a.mem_v = att.mem_v;
            a.n_mem = m.n_memory;
        }
        a.B = B; a.H = m.heads; a.nq = n; a.nk = n;
        a.scale = 1.0f / std::sqrt(static_cast<float>(m.d_k));
        return cap_attention(&a, s);
    };
    if (e->enc_chains && !(dbg_ablate() & 2)) {
        // the encoder as chains: 1 + layers launches of the chain kernel with the self-attention kernel between them;
        // the cross K|V projections ride on the stages whose level output they read
        CAP_PROPAGATE(cap_enc_chain(e->enc_chains, 0, rows, n, s));
        for (int l = 0; l < m.enc_layers; ++l) {
            CAP_PROPAGATE(self_attention(e->enc[l].att));
            CAP_PROPAGATE(cap_enc_chain(e->enc_chains, 1 + l, rows, n, s));
        }
        e->encoded = true;
        return CAP_OK;
    }
    // one kernel per operator.  E0: LN(projection) + pos
    CAP_PROPAGATE(run_linear_ln(e, e->feat_bf16, m.d_feature, e->vis_proj, nullptr, e->enc_ln, e->vis_pos, n, nullptr,
                                e->buf_x, e->res_x, rows, s));

    const bf16* x = e->buf_x;
    for (int l = 0; l < ((dbg_ablate() & 2) ? 0 : m.enc_layers); ++l) {
        const EncoderLayerW& L = e->enc[l];
        bf16* level_out = e->enc_levels + static_cast<size_t>(l) * rows_cap * d;
        CAP_PROPAGATE(run_linear(x, d, L.att.qkv, e->buf_qkv, 3 * hd, CAP_BF16, CAP_ACT_NONE, rows, s));
        CAP_PROPAGATE(self_attention(L.att));
        CAP_PROPAGATE(run_linear_ln(e, e->buf_att, hd, L.att.o, e->res_x, L.att.ln, nullptr, 0, nullptr, e->buf_a, e->res_a,
                                    rows, s));
        if (m.aoa_enc) CAP_PROPAGATE(run_aoa(e, L.att, x, e->buf_a, e->buf_a, e->res_a, rows, s));
        CAP_PROPAGATE(run_linear(e->buf_a, d, L.ffn.fc1, e->buf_h, m.d_ff, CAP_BF16, CAP_ACT_RELU, rows, s));
        CAP_PROPAGATE(run_linear_ln(e, e->buf_h, m.d_ff, L.ffn.fc2, e->res_a, L.ffn.ln, nullptr, 0, e->enc_mask, level_out,
                                    e->res_x, rows, s));
        x = level_out;
    }
    // Cross-attention K/V, once per image per decoder layer (and per level for the meshed decoder).
    for (int l = 0; l < m.dec_layers; ++l) {
        for (int i = 0; i < lv; ++i) {
            const int src_level = (m.decoder_kind == CAP_DEC_MESHED) ? i : m.enc_layers - 1;
            const bf16* src = e->enc_levels + static_cast<size_t>(src_level) * rows_cap * d;
            bf16* dst = e->cross_kv + (static_cast<size_t>(l) * lv + i) * rows_cap * 2 * hd;
            CAP_PROPAGATE(run_linear(src, d, e->dec[l].cross_att.kv, dst, 2 * hd, CAP_BF16, CAP_ACT_NONE, rows, s));
        }
    }
    e->encoded = true;
    return CAP_OK;
}

extern "C" int cap_engine_begin_decode(cap_engine* e, cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_begin_decode: encode first");
    return cap_beam_reset(e->beam_state, e->cur_batch, e->desc.bos_idx, stream);
}

namespace {
int run_decoder_stack(cap_engine* e, int t, cudaStream_t s, bf16** hidden);


// Decoder stack + vocabulary projection (logits and chunk statistics) of step t: GEMM chains with the attention
// kernels between them (decoders.py:21-28 / 51-73 per layer).
int run_fused_stack(cap_engine* e, int t, cudaStream_t s) {
    const cap_model_desc& m = e->desc;
    const int hd = e->hd(), T = m.max_len, B = e->cur_batch, R = B * e->beam, lv = e->levels();
    const size_t rows_cap = static_cast<size_t>(e->max_batch) * e->n_tokens;
    const size_t max_rows = static_cast<size_t>(e->max_batch) * e->beam;
    const float scale = 1.0f / std::sqrt(static_cast<float>(m.d_k));
    const int ab = dbg_ablate();
    if (!(ab & 4)) CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_EMBED_QKV, 0, t, B, s));
    for (int l = 0; l < m.dec_layers; ++l) {
        const bf16* cache_l = e->qkv_cache + static_cast<size_t>(l) * T * R * 3 * hd;
        if (!(ab & 1))
            CAP_PROPAGATE(cap_decode_self_attention(cache_l, cap_beam_ancestry(e->beam_state), e->padflag, e->buf_att, hd, t,
                                                    R, m.heads, scale, s));
        if (!(ab & 4)) CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_SELF_OUT, l, t, B, s));
        if (!(ab & 8)) {
            // every encoder level in ONE launch: level i reads cross_kv[l][i], writes att_in[i] (same queries)
            const bf16* kv = e->cross_kv + static_cast<size_t>(l) * lv * rows_cap * 2 * hd;
            CAP_PROPAGATE(cap_decode_cross_attention_levels(e->buf_q, hd, kv, rows_cap * 2 * hd, e->enc_mask, e->buf_att, hd,
                                                            max_rows * hd, B, e->beam, e->cur_n, m.heads, lv, scale, s));
        }
        if (!(ab & 4)) CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_FFN, l, t, B, s));
    }
    return CAP_OK;
}
}

extern "C" int cap_engine_decode_logits(cap_engine* e, int t, cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_decode_logits: encode first");
    CAP_REQUIRE(t >= 0 && t < e->desc.max_len, "cap_engine_decode_logits: step %d outside [0,%d)", t, e->desc.max_len);
    bf16* x = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->fused) {  // the production stack; this debug entry point wants EVERY logit stored
        const int was = cap_fused_get_full_logits(e->fused);
        cap_fused_set_full_logits(e->fused, 1);
        const int rc = run_fused_stack(e, t, s);
        cap_fused_set_full_logits(e->fused, was);
        return rc;
    }
    CAP_PROPAGATE(run_decoder_stack(e, t, s, &x));
    // bias-free vocabulary projection (decoders.py:90,121); log-softmax happens in the beam row pass
    return run_linear(x, e->desc.d_model, e->vocab_fc, e->logits, e->ld_logits, CAP_F32, CAP_ACT_NONE,
                      e->cur_batch * e->beam, s);
}

extern "C" int cap_engine_decode_step(cap_engine* e, int t, cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_decode_step: encode first");
    CAP_REQUIRE(t >= 0 && t < e->desc.max_len, "cap_engine_decode_step: step %d outside [0,%d)", t, e->desc.max_len);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->vocab_chunks > 512)  // vocabularies beyond the merge kernel's reach: full row pass over the logits
    {
        CAP_PROPAGATE(cap_engine_decode_logits(e, t, stream));
        return cap_engine_beam_advance(e, t, stream);
    }
    if (e->fused) {
        CAP_PROPAGATE(run_fused_stack(e, t, s));
        return cap_beam_step_stats(e->beam_state, t, e->logits, e->ld_logits, e->part_ms, e->vocab_chunks, s);
    }
    bf16* x = nullptr;
    CAP_PROPAGATE(run_decoder_stack(e, t, s, &x));
    const int R = e->cur_batch * e->beam;
    int chunks = 0;
    CAP_PROPAGATE(cap_vocab_logits_stats(x, e->desc.d_model, e->vocab_fc.w, e->vocab_fc.b, e->logits, e->ld_logits, R,
                                         e->desc.vocab, e->desc.d_model, e->part_ms, &chunks, s));
    return cap_beam_step_stats(e->beam_state, t, e->logits, e->ld_logits, e->part_ms, chunks, s);
}

namespace {
int run_decoder_stack(cap_engine* e, int t, cudaStream_t s, bf16** hidden) {
    const cap_model_desc& m = e->desc;
    const int d = m.d_model, hd = e->hd(), lv = e->levels(), T = m.max_len;
    const int B = e->cur_batch, n = e->cur_n, beam = e->beam;
    const int R = B * beam;
    const size_t rows_cap = static_cast<size_t>(e->max_batch) * e->n_tokens;
    const float scale = 1.0f / std::sqrt(static_cast<float>(m.d_k));
    uint8_t* pad_t = e->padflag + static_cast<size_t>(t) * R;

    // D3: x = Emb[token] + pos[t+1]   (running_seq is t+1 for every row, decoders.py:107-109)
    CAP_PROPAGATE(cap_embed_tokens(cap_beam_tokens(e->beam_state), e->word_emb, e->word_pos, t + 1, m.pad_idx, e->buf_x,
                                   e->res_x, pad_t, R, d, s));
    bf16* x = e->buf_x;
    for (int l = 0; l < m.dec_layers; ++l) {
        const DecoderLayerW& L = e->dec[l];
        bf16* cache_l = e->qkv_cache + static_cast<size_t>(l) * T * R * 3 * hd;
        // A5 (self): project q|k|v of the new token straight into the cache slot of step t
        CAP_PROPAGATE(run_linear(x, d, L.self_att.qkv, cache_l + static_cast<size_t>(t) * R * 3 * hd, 3 * hd, CAP_BF16,
                                 CAP_ACT_NONE, R, s));
        CAP_PROPAGATE(cap_decode_self_attention(cache_l, cap_beam_ancestry(e->beam_state), e->padflag, e->buf_att, hd,
                                                t, R, m.heads, scale, s));
        CAP_PROPAGATE(run_linear_ln(e, e->buf_att, hd, L.self_att.o, e->res_x, L.self_att.ln, nullptr, 0, nullptr, e->buf_a,
                                    e->res_a, R, s));
        if (m.aoa_dec_self) CAP_PROPAGATE(run_aoa(e, L.self_att, x, e->buf_a, e->buf_a, e->res_a, R, s));
        const bf16* sa = e->buf_a;
        // A5 (cross): one q projection shared by every encoder level (same enc_attn weights)
        CAP_PROPAGATE(run_linear(sa, d, L.cross_att.q, e->buf_q, hd, CAP_BF16, CAP_ACT_NONE, R, s));
        for (int i = 0; i < lv; ++i) {
            const bf16* kv = e->cross_kv + (static_cast<size_t>(l) * lv + i) * rows_cap * 2 * hd;
            CAP_PROPAGATE(cap_decode_cross_attention(e->buf_q, hd, kv, e->enc_mask, e->buf_att + static_cast<size_t>(i) * R * hd,
                                                     hd, B, beam, n, m.heads, scale, s));
        }
        if (!(e->fuse_ln && d % 128 == 0 && d <= 1024))
            CAP_PROPAGATE(run_linear(e->buf_att, hd, L.cross_att.o, e->buf_y32, d, CAP_F32, CAP_ACT_NONE, lv * R, s));
        for (int i = 0; i < lv; ++i) {
            bf16* ci = e->buf_c + static_cast<size_t>(i) * R * d;
            float* ci32 = e->res_c + static_cast<size_t>(i) * R * d;
            if (e->fuse_ln && d % 128 == 0 && d <= 1024)  // every level adds the same residual (decoders.py:56)
                CAP_PROPAGATE(run_linear_ln(e, e->buf_att + static_cast<size_t>(i) * R * hd, hd, L.cross_att.o, e->res_a,
                                            L.cross_att.ln, nullptr, 0, nullptr, ci, ci32, R, s));
            else
                CAP_PROPAGATE(run_ln(e->buf_y32 + static_cast<size_t>(i) * R * d, e->res_a, L.cross_att.ln, nullptr, 0,
                                     nullptr, ci, ci32, R, d, s));
            if (m.aoa_dec_cross) CAP_PROPAGATE(run_aoa(e, L.cross_att, sa, ci, ci, ci32, R, s));
        }
        const bf16* c = e->buf_c;
        const float* c32 = e->res_c;
        if (m.decoder_kind == CAP_DEC_MESHED) {
            // D2: alpha_i = sigmoid(W_i [s ; c_i]),  c = sum_i alpha_i * c_i / sqrt(levels)
            float* gates = e->buf_y32;
            for (int i = 0; i < lv; ++i) {
                const bf16* ci = e->buf_c + static_cast<size_t>(i) * R * d;
                CAP_CHECK_CUDA(cudaMemcpy2DAsync(e->buf_cat, 2 * d * 2, sa, d * 2, d * 2, R, cudaMemcpyDeviceToDevice, s));
                CAP_CHECK_CUDA(cudaMemcpy2DAsync(e->buf_cat + d, 2 * d * 2, ci, d * 2, d * 2, R, cudaMemcpyDeviceToDevice, s));
                CAP_PROPAGATE(run_linear(e->buf_cat, 2 * d, L.alphas[i], gates + static_cast<size_t>(i) * R * d, d, CAP_F32,
                                         CAP_ACT_NONE, R, s));
            }
            CAP_PROPAGATE(cap_meshed_mix(gates, e->res_c, CAP_F32, e->buf_mix, e->res_mix, lv, R, d, s));
            c = e->buf_mix;
            c32 = e->res_mix;
        }
        // F1 + zero rows whose input token was <pad> (decoders.py:26)
        CAP_PROPAGATE(run_linear(c, d, L.ffn.fc1, e->buf_h, m.d_ff, CAP_BF16, CAP_ACT_RELU, R, s));
        CAP_PROPAGATE(run_linear_ln(e, e->buf_h, m.d_ff, L.ffn.fc2, c32, L.ffn.ln, nullptr, 0, pad_t, e->buf_x, e->res_x, R, s));
        x = e->buf_x;
    }
    *hidden = x;
    return CAP_OK;
}
}  // namespace

extern "C" int cap_engine_beam_advance(cap_engine* e, int t, cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_beam_advance: encode first");
    return cap_beam_step(e->beam_state, t, e->logits, e->ld_logits, 0, stream);
}

namespace {
int run_search_eager(cap_engine* e, int out_size, int64_t* ids, float* logp, cudaStream_t s) {
    CAP_PROPAGATE(cap_engine_begin_decode(e, s));
    for (int t = 0; t < e->desc.max_len; ++t) CAP_PROPAGATE(cap_engine_decode_step(e, t, s));
    return cap_beam_finalize(e->beam_state, out_size, ids, logp, s);
}
}  // namespace

extern "C" int cap_engine_beam_search(cap_engine* e, int out_size, int64_t* ids, float* logp, int use_graph,
                                      cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_beam_search: encode first");
    CAP_REQUIRE(ids && logp, "cap_engine_beam_search: null output");
    CAP_REQUIRE(out_size >= 1 && out_size <= e->beam, "cap_engine_beam_search: out_size outside [1,beam]");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!use_graph || !e->warmed) {
        // the first run is always eager: it sets kernel attributes, which must not happen mid-capture
        e->warmed = true;
        return run_search_eager(e, out_size, ids, logp, s);
    }
    const bool hit = e->graph_exec && e->graph_batch == e->cur_batch && e->graph_n == e->cur_n &&
                     e->graph_out_size == out_size && e->graph_ids == ids && e->graph_logp == logp;
    if (!hit) {
        if (e->graph_exec) {
            cudaGraphExecDestroy(e->graph_exec);
            e->graph_exec = nullptr;
        }
        cudaGraph_t graph = nullptr;
        if (!e->capture_stream) CAP_CHECK_CUDA(cudaStreamCreateWithFlags(&e->capture_stream, cudaStreamNonBlocking));
        cudaStream_t cs = e->capture_stream;
        CAP_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        const int rc = run_search_eager(e, out_size, ids, logp, cs);
        const cudaError_t end = cudaStreamEndCapture(cs, &graph);
        if (rc != CAP_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (end != cudaSuccess) return cap_set_error(CAP_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(end));
        const cudaError_t inst = cudaGraphInstantiate(&e->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (inst != cudaSuccess) return cap_set_error(CAP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(inst));
        e->graph_batch = e->cur_batch;
        e->graph_n = e->cur_n;
        e->graph_out_size = out_size;
        e->graph_ids = ids;
        e->graph_logp = logp;
    }
    CAP_CHECK_CUDA(cudaGraphLaunch(e->graph_exec, s));
    return CAP_OK;
}

extern "C" int cap_engine_caption_host_async(cap_engine* e, const void* feats_host, int feat_dtype,
                                             const float* boxes_host, int B, int n, int out_size, int64_t* ids_host,
                                             float* logp_host, int use_graph, cap_stream_t stream);

extern "C" int cap_engine_caption_host(cap_engine* e, const void* feats_host, int feat_dtype, const float* boxes_host,
                                       int B, int n, int out_size, int64_t* ids_host, float* logp_host, int use_graph,
                                       cap_stream_t stream) {
    CAP_PROPAGATE(cap_engine_caption_host_async(e, feats_host, feat_dtype, boxes_host, B, n, out_size, ids_host,
                                                logp_host, use_graph, stream));
    CAP_CHECK_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return CAP_OK;
}

extern "C" int cap_engine_caption_host_async(cap_engine* e, const void* feats_host, int feat_dtype,
                                             const float* boxes_host, int B, int n, int out_size, int64_t* ids_host,
                                             float* logp_host, int use_graph, cap_stream_t stream) {
    CAP_REQUIRE(e && e->max_batch > 0, "cap_engine_caption_host: reserve the engine first");
    CAP_REQUIRE(feats_host && ids_host && logp_host, "cap_engine_caption_host: null pointer");
    CAP_REQUIRE(B > 0 && B <= e->max_batch && n > 0 && n <= e->n_tokens, "cap_engine_caption_host: bad batch shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const cap_model_desc& m = e->desc;
    const size_t esz = feat_dtype == CAP_F32 ? 4 : 2;
    const size_t feat_bytes = static_cast<size_t>(B) * n * m.d_feature * esz;
    CAP_CHECK_CUDA(cudaMemcpyAsync(e->feat_stage, feats_host, feat_bytes, cudaMemcpyHostToDevice, s));
    const float* boxes_dev = nullptr;
    if (boxes_host) {
        CAP_CHECK_CUDA(cudaMemcpyAsync(e->box_stage, boxes_host, static_cast<size_t>(B) * n * 16, cudaMemcpyHostToDevice, s));
        boxes_dev = e->box_stage;
    }
    CAP_PROPAGATE(cap_engine_encode(e, e->feat_stage, feat_dtype, boxes_dev, B, n, s));
    CAP_PROPAGATE(cap_engine_beam_search(e, out_size, e->out_ids, e->out_logp, use_graph, s));
    const size_t count = static_cast<size_t>(B) * out_size * m.max_len;
    CAP_CHECK_CUDA(cudaMemcpyAsync(ids_host, e->out_ids, count * 8, cudaMemcpyDeviceToHost, s));
    CAP_CHECK_CUDA(cudaMemcpyAsync(logp_host, e->out_logp, count * 4, cudaMemcpyDeviceToHost, s));
    return CAP_OK;
}

// Device-resident variant of the call above: features already in HBM, results left in HBM (engine-owned
// decode + one small device-to-device copy into the caller's buffers), nothing synchronises.
extern "C" int cap_engine_caption_device_async(cap_engine* e, const void* feats_dev, int feat_dtype, const float* boxes_dev,
                                               int B, int n, int out_size, int64_t* ids_dev, float* logp_dev, int use_graph,
                                               cap_stream_t stream) {
    CAP_REQUIRE(e && e->max_batch > 0, "cap_engine_caption_device: reserve the engine first");
    CAP_REQUIRE(feats_dev && ids_dev && logp_dev, "cap_engine_caption_device: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CAP_PROPAGATE(cap_engine_encode(e, feats_dev, feat_dtype, boxes_dev, B, n, s));
    CAP_PROPAGATE(cap_engine_beam_search(e, out_size, e->out_ids, e->out_logp, use_graph, s));
    const size_t count = static_cast<size_t>(B) * out_size * e->desc.max_len;
    CAP_CHECK_CUDA(cudaMemcpyAsync(ids_dev, e->out_ids, count * 8, cudaMemcpyDeviceToDevice, s));
    CAP_CHECK_CUDA(cudaMemcpyAsync(logp_dev, e->out_logp, count * 4, cudaMemcpyDeviceToDevice, s));
    return CAP_OK;
}

// Measurement hook (bench.py roofline): the GEMM chains of step t back to back WITHOUT the attention and beam
// kernels between them -- same launches, same shapes, same weights as in a real step (the activations they read
// are whatever the last real step left behind).  Returns CAP_ERR_STATE when the engine does not run chains.
extern "C" int cap_engine_debug_chains(cap_engine* e, int t, cap_stream_t stream) {
    CAP_REQUIRE(e && e->encoded, "cap_engine_debug_chains: encode first");
    if (!e->fused) return cap_set_error(CAP_ERR_STATE, "cap_engine_debug_chains: engine is not in chain mode");
    CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_EMBED_QKV, 0, t, e->cur_batch, stream));
    for (int l = 0; l < e->desc.dec_layers; ++l) {
        CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_SELF_OUT, l, t, e->cur_batch, stream));
        CAP_PROPAGATE(cap_fused_chain(e->fused, CAP_CHAIN_FFN, l, t, e->cur_batch, stream));
    }
    return CAP_OK;
}

extern "C" const void* cap_engine_encoder_output(cap_engine* e) { return e ? e->enc_levels : nullptr; }
extern "C" const uint8_t* cap_engine_encoder_mask(cap_engine* e) { return e ? e->enc_mask : nullptr; }
extern "C" const float* cap_engine_logits(cap_engine* e, int* ld) {
    if (!e) return nullptr;
    if (ld) *ld = e->ld_logits;
    return e->logits;
}
extern "C" cap_beam* cap_engine_beam(cap_engine* e) { return e ? e->beam_state : nullptr; }
