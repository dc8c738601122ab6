// QWen3.cpp -- Qwen3 decode / prefill on top of the device C ABI (see QWen3.hpp for the reference interfaces mirrored here).
#include "QWen3.hpp"

#include "../Tensor/KunFile.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace koifish {

#define KF_TRY(expr)                                                                                   \
    do {                                                                                               \
        int _rc = (expr);                                                                              \
        if (_rc != KF_OK) {                                                                            \
            if (hFishErr) *hFishErr = std::string(#expr) + " -> " + kf_status_string(_rc);             \
            return _rc;                                                                                \
        }                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------------------ config
static int jint(const JSON& j, std::initializer_list<const char*> path, int dflt) {
    const JSON* p = j.path(path);
    return p ? p->as_int(dflt) : dflt;
}
MODEL_CARD MODEL_CARD::FromJSON(const JSON& j0) {
    MODEL_CARD c;
    const JSON* hf = j0.find("hf_config");
    if (!hf && j0.contains("hidden_size")) hf = &j0;
    if (hf) {  // HF config.json (MODEL_CARD::InitHugFace, src/Utils/CLI_params.cpp:2224-2300)
        const JSON& h = *hf;
        c.n_embd    = jint(h, {"hidden_size"}, 0);
        c.n_ff      = jint(h, {"intermediate_size"}, 0);
        c.n_layers  = jint(h, {"num_hidden_layers"}, 0);
        c.n_head    = jint(h, {"num_attention_heads"}, 0);
        c.n_head_kv = jint(h, {"num_key_value_heads"}, c.n_head);
        c.head_dim  = jint(h, {"head_dim"}, c.n_head ? c.n_embd / c.n_head : 128);
        c.vocab     = jint(h, {"vocab_size"}, c.vocab);
        c.max_pos_embeddings = jint(h, {"max_position_embeddings"}, c.max_pos_embeddings);
        if (const JSON* t = h.find("rope_theta")) c.rope_theta = (float)t->as_double(c.rope_theta);
        if (const JSON* e = h.find("rms_norm_eps")) c.norm_rms_eps = (float)e->as_double(c.norm_rms_eps);
        if (const JSON* t = h.find("tie_word_embeddings")) c.tie_word_embeddings = t->as_bool(false);
        if (const JSON* vq = h.find("quantization_config")) {
            // QUANT_CARD::Vendor2JSONx (src/Utils/CLI_params.cpp:240-262, called from MODEL_CARD::InitHugFace :2285-2290): the vendor's block
            // {"bits", "group_size", "quant_method": "awq", "zero_point", ...} becomes the card of every self_attn / mlp linear
            if (!vq->is_object() || vq->empty()) throw std::runtime_error("HF quantization_config must be a non-empty object");
            JSON q, yes;
            q.kind = JSON::Object, yes.kind = JSON::Bool, yes.b = true;
            q.obj.push_back({"self_attn", *vq});
            q.obj.push_back({"mlp", *vq});
            q.obj.push_back({"VendorQuant", yes});
            if (vq->contains("quant_method") && vq->at("quant_method").as_string() == "awq") q.obj.push_back({"ExplicitZS", yes});
            c.jQuant = q;
        }
    }
    if (const JSON* m = j0.find("model")) {  // Koifish JSON (cases/qwen3/*.json)
        if (const JSON* a = m->find("arch")) c.arch = a->as_string(c.arch);
        c.n_layers  = jint(*m, {"parameter", "Layer"}, c.n_layers);
        c.n_ctx     = jint(*m, {"parameter", "transformer", "Ctx"}, c.n_ctx);
        c.n_embd    = jint(*m, {"parameter", "transformer", "Embed"}, c.n_embd);
        c.n_ff      = jint(*m, {"parameter", "transformer", "Ffn"}, c.n_ff);
        c.n_head    = jint(*m, {"parameter", "transformer", "Head"}, c.n_head);
        c.n_head_kv = jint(*m, {"parameter", "transformer", "KVHead"}, c.n_head_kv ? c.n_head_kv : c.n_head);
        c.head_dim  = jint(*m, {"parameter", "transformer", "head_dim"}, c.head_dim);
        c.vocab     = jint(*m, {"parameter", "vocab_size"}, c.vocab);
        c.max_pos_embeddings = jint(*m, {"parameter", "max_pos_embeddings"}, c.max_pos_embeddings);
        if (const JSON* t = m->path({"parameter", "tie_word_embeddings"})) c.tie_word_embeddings = t->as_bool(false);
        if (const JSON* t = m->path({"parameter", "rope_theta"})) c.rope_theta = (float)t->as_double(c.rope_theta);
    }
    if (const JSON* q = j0.find("quantizer")) c.jQuant = *q;  // "# quantizer" (commented out) is simply another key
    c.seed        = jint(j0, {"seed"}, c.seed);
    c.max_seq_len = jint(j0, {"gpt", "max_seq_len"}, c.max_seq_len);
    c.max_batch   = jint(j0, {"gpt", "max_batch"}, c.max_batch);
    c.max_prefill = jint(j0, {"gpt", "max_prefill"}, c.max_prefill);
    c.awq_repack  = jint(j0, {"gpt", "awq_repack"}, c.awq_repack);
    if (const JSON* s = j0.path({"init", "sigma"})) c.init_sigma = (float)s->as_double(c.init_sigma);
    if (const JSON* s = j0.path({"init", "norm_sigma"})) c.norm_sigma = (float)s->as_double(c.norm_sigma);
    std::string a = c.arch;
    std::transform(a.begin(), a.end(), a.begin(), ::toupper);
    if (a != "QWEN3") throw std::runtime_error("model.arch '" + c.arch + "' is outside the hot path (only QWEN3 is built)");
    if (c.n_layers <= 0 || c.n_embd <= 0 || c.n_ff <= 0 || c.n_head <= 0 || c.n_head_kv <= 0 || c.head_dim <= 0 || c.vocab <= 0)
        throw std::runtime_error("model config incomplete: need Layer/Embed/Ffn/Head/KVHead/head_dim");
    if (c.n_head % c.n_head_kv) throw std::runtime_error("Head must be a multiple of KVHead");
    if (c.head_dim != 64 && c.head_dim != 128) throw std::runtime_error("head_dim must be 64 or 128");
    if (c.max_seq_len <= 0 || c.max_batch <= 0) throw std::runtime_error("gpt.max_seq_len / gpt.max_batch must be positive");
    return c;
}

// ------------------------------------------------------------------------------------------------------------ neurons
void* KVCache::Get(CTYPE t, int layer, int pos, int seq) const {
    uint16_t* base = (uint16_t*)(t == KV_KEY ? key : value);
    return base + (((size_t)layer * max_batch + seq) * max_seq + pos) * kv_dim;
}
int SLP::Forw(void* rhs, const void* lhs, int M, int epilogue, const void* residual) {
    kf_tensor_desc d = w->Desc();
    return kf_linear(hFish->ctx, rhs, &d, lhs, M, epilogue, residual);
}
int LayerNormal::cuFlow(void* out, const void* inp, int rows) { return kf_rmsnorm(hFish->ctx, out, inp, w->data, rows, w->ne[1], rms_eps); }

int ROPE::cuInfer(SelfAttention* a, int M) {
    Fish* f        = hFish;
    const int lay  = a->layid - 1;
    const size_t ss = f->seq_mode ? f->cache.seq_stride() : 0;
    return kf_qknorm_rope_kvappend(f->ctx, f->q, f->k, f->v, q_norm ? q_norm->data : nullptr, k_norm ? k_norm->data : nullptr,
                                   f->cache.Get(KVCache::KV_KEY, lay), f->cache.Get(KVCache::KV_VAL, lay), table, f->d_pos, M, a->n_head,
                                   a->n_head_kv, a->head_dim, f->cache.max_seq, 1e-6f, ss);
}

// SelfAttention::cuInfer, reference src/Device/CUDA/QKV.cu:617-706:
//   norm -> Q/K/V.Forw -> rope->cuInfer (QK-norm, rope, K/V into the cache) -> attention -> proj_cat.Forw -> residual add
int SelfAttention::cuInfer(void* inpL, int M) {
    Fish* f = hFish;
    std::string* hFishErr = &f->error;
    kf_tensor_desc w3[3] = {Q.w->Desc(), K.w->Desc(), V.w->Desc()};
    void* y3[3]          = {f->q, f->k, f->v};
    // norm.cuFlow + the three SLP::Forw calls share one launch; the normalised activations never leave the chip
    KF_TRY(kf_rmsnorm_linear(f->ctx, 3, y3, w3, inpL, norm.w->data, norm.rms_eps, M, 0));
    const int lay   = layid - 1;
    const size_t ss = f->seq_mode ? f->cache.seq_stride() : 0;
    int gqa_min_ctx = 1024;
    kf_ctx_get_int(f->ctx, "gqa_min_ctx", &gqa_min_ctx);
    const bool long_ctx = M == 1 && f->attn_hint + 1 > gqa_min_ctx;  // one sequence, long context: stream each cached row once per kv head
    // several sequences: the fused per-head kernel stays resident in one wave up to 64 (token, head) pairs; beyond that the kv-group
    // kernel wins (measured: Qwen3-32B from 2 sequences, Qwen3-8B from 3)
    const bool many = f->seq_mode && (f->gqa_min_batch > 0 ? M >= f->gqa_min_batch : n_head * M > 64);
    if ((many || long_ctx) && n_head / n_head_kv <= 16) {
        // many sequences: QK-norm + RoPE + append, then the kv-group attention on the tensor cores (each cached row read once per kv head)
        KF_TRY(rope.cuInfer(this, M));
        KF_TRY(kf_attn_decode_gqa(f->ctx, f->att, f->q, f->cache.Get(KVCache::KV_KEY, lay), f->cache.Get(KVCache::KV_VAL, lay), f->d_pos, M, n_head,
                                  n_head_kv, head_dim, f->cache.max_seq, f->attn_hint, ss));
    } else if (M == 1 || f->seq_mode) {  // decode: rope->cuInfer + the attention kernels in one launch
        KF_TRY(kf_qkv_attention(f->ctx, f->att, f->q, f->k, f->v, rope.q_norm ? rope.q_norm->data : nullptr,
                                rope.k_norm ? rope.k_norm->data : nullptr, f->cache.Get(KVCache::KV_KEY, lay), f->cache.Get(KVCache::KV_VAL, lay),
                                rope.table, f->d_pos, M, n_head, n_head_kv, head_dim, f->cache.max_seq, 1e-6f, ss, f->attn_hint));
    } else {  // prefill panel: the tokens attend to each other's fresh K/V rows, so the append must complete first
        KF_TRY(rope.cuInfer(this, M));
        if (M >= 16 && f->panel_consecutive)  // tensor-core flash attention over the panel (positions pos[0] .. pos[0] + M - 1)
            KF_TRY(kf_attn_prefill(f->ctx, f->att, f->q, f->cache.Get(KVCache::KV_KEY, lay), f->cache.Get(KVCache::KV_VAL, lay), f->d_pos, M, n_head,
                                   n_head_kv, head_dim, f->cache.max_seq));
        else
            KF_TRY(kf_attn_decode(f->ctx, f->att, f->q, f->cache.Get(KVCache::KV_KEY, lay), f->cache.Get(KVCache::KV_VAL, lay), f->d_pos, M, n_head,
                                  n_head_kv, head_dim, f->cache.max_seq, f->attn_hint, ss));
    }
    const size_t nE = (size_t)M * f->config.n_embd;
    if (f->tp_world == 1) {
        KF_TRY(proj_cat.Forw(inpL, f->att, M, KF_EPI_RESIDUAL, inpL));  // out = residual + proj (CU_add3, QKV.cu:682-688)
    } else if (f->tp_fuse) {  // row-parallel, the exchange + residual add in the epilogue of the matmul (kf_tp.cuh): no launch of its own
        kf_tensor_desc d = proj_cat.w->Desc();
        KF_TRY(kf_linear_exchange(f->ctx, inpL, &d, f->att, M, inpL));
    } else {  // row-parallel: fp32 partial sums -> all-reduce over NVLink -> residual add
        KF_TRY(proj_cat.Forw(f->part_f32, f->att, M, KF_EPI_F32, nullptr));
        KF_TRY(kf_allreduce_residual(f->ctx, inpL, inpL, f->part_f32, nE));  // one launch over NVLink peer memory (p2p.cu); NCCL fallback
    }
    return KF_OK;
}
// FFN::cuInfer, reference src/Device/CUDA/NeuronFuse.cu:615-656: norm -> gate.Forw, up.Forw -> relu.Forw (SwiGLU) -> down.Forw -> add
int FFN::cuInfer(void* inpL, int M) {
    Fish* f = hFish;
    std::string* hFishErr = &f->error;
    kf_tensor_desc gu[2] = {gate.w->Desc(), up.w->Desc()};
    void* y2[2]          = {f->hb, f->hb};
    KF_TRY(kf_rmsnorm_linear(f->ctx, 2, y2, gu, inpL, norm.w->data, norm.rms_eps, M, 2));  // norm + gate/up + SwiGLU in one launch
    const size_t nE = (size_t)M * f->config.n_embd;
    if (f->tp_world == 1) {
        KF_TRY(down.Forw(inpL, f->hb, M, KF_EPI_RESIDUAL, inpL));
    } else if (f->tp_fuse) {
        kf_tensor_desc d = down.w->Desc();
        KF_TRY(kf_linear_exchange(f->ctx, inpL, &d, f->hb, M, inpL));
    } else {
        KF_TRY(down.Forw(f->part_f32, f->hb, M, KF_EPI_F32, nullptr));
        KF_TRY(kf_allreduce_residual(f->ctx, inpL, inpL, f->part_f32, nE));  // one launch over NVLink peer memory (p2p.cu); NCCL fallback
    }
    return KF_OK;
}
int TokenEmbed::cuInfer(void* out, int M) {
    kf_tensor_desc d = w->Desc();
    return kf_embed(hFish->ctx, out, &d, hFish->d_tokens, M);
}
// Head4Token::cuInfer_1 (NeuronFuse.cu:842-862) preceded by the final LayerNormal.  Under tensor parallelism the vocabulary rows are
// sharded and the logits all-gathered.
int Head4Token::cuInfer_1(void* logits, const void* inp, int M) {
    Fish* f = hFish;
    std::string* hFishErr = &f->error;
    kf_tensor_desc d = proj.w->Desc();
    const int W = f->tp_world;
    if (W == 1) {
        void* y1[1] = {logits};
        KF_TRY(kf_rmsnorm_linear(f->ctx, 1, y1, &d, inp, norm.w->data, norm.rms_eps, M, 0));  // final norm folded into the lm_head GEMV
        return KF_OK;
    }
    KF_TRY(norm.cuFlow(f->xb, inp, M));
    const int vl = f->config.vocab / W;
    const typNUMBER tp = proj.w->type;
    const size_t row_bytes = (size_t)((double)d.cols * BitPE(tp) / 8);
    d.data_dev = (const uint8_t*)d.data_dev + (size_t)f->tp_rank * vl * row_bytes;  // row-slice view of the (replicated) table
    if (d.gama_dev) {
        const int gpr      = d.cols / d.group;
        const uint16_t* g0 = (const uint16_t*)d.gama_dev + d.rows + d.cols;
        d.zero_dev = g0 + (size_t)f->tp_rank * vl * gpr;
        d.step_dev = g0 + (size_t)d.rows * gpr + (size_t)f->tp_rank * vl * gpr;
    }
    d.rows = vl;
    // local logits land in the tail of the buffer, then all ranks' pieces are gathered to the front: [W][M][vl]
    uint16_t* local = (uint16_t*)logits + (size_t)f->logit_rows * f->config.vocab;
    KF_TRY(kf_linear(f->ctx, local, &d, f->xb, M, KF_EPI_NONE, nullptr));
    KF_TRY(kf_allgather(f->ctx, logits, local, (size_t)M * vl * 2));
    if (M > 1) {  // [W][M][vl] -> [M][W*vl]: one copy out of the way + one re-layout kernel (not W x M memcpy nodes)
        uint16_t* tmp = local;
        KF_TRY(kf_d2d(f->ctx, tmp, logits, (size_t)M * f->config.vocab * 2));
        KF_TRY(kf_relayout_wmv(f->ctx, logits, tmp, W, M, vl));
    }
    return KF_OK;
}

// ------------------------------------------------------------------------------------------------------------ model
Fish::Fish(kf_ctx* c, const MODEL_CARD& card, int rank, int world) : ctx(c), config(card), tp_rank(rank), tp_world(world) {}
Fish::~Fish() {
    ResetGraphs();
    tensors.clear();
    void* bufs[] = {x, xb, q, k, v, att, hb, logits, part_f32, d_tokens, d_pos, d_next, d_rng, cache.key, cache.value};
    for (void* b : bufs)
        if (b) kf_free(ctx, b);
    if (rope_table_shared) kf_free(ctx, rope_table_shared);
    if (h_stage) kf_host_free(h_stage);
    if (h_logits) kf_host_free(h_logits);
}
void Fish::ResetGraphs() {
    weights_dirty = true;  // every path that changes a resident tensor ends here: kf_model_info_get recounts the resident bytes once
    for (auto& g : graphs)
        if (g.second) kf_graph_destroy(g.second);
    graphs.clear();
}
// A captured graph holds raw pointers into the context's scratch buffers (split-K / attention workspaces, tensor-core staging).  Those
// buffers are re-allocated when a later eager call -- a bigger batch, a prefill panel, another model on the same context -- needs more
// room; the context counts such events and every replay path checks the count first.
void Fish::SyncGraphGeneration() {
    const uint64_t gen = kf_scratch_generation(ctx);
    if (gen != graph_gen) {
        ResetGraphs();
        graph_gen = gen;
    }
}
int Fish::AllocTensor(const std::string& name, int rows, int cols, int id, hGTensor& out) {
    out = std::make_shared<GTensor>(ctx, name, rows, cols);
    // only 2-D weight matrices are quantised, norms never (isWMAT, GeQuant.cpp:155-156)
    if (rows > 1) out->hQuant = GeQuant::MakeInstance(name, config.jQuant);
    tensors[name]    = out;
    tensor_ids[name] = id;
    return KF_OK;
}

// Fish::MakeInstance -> QWen3 ctor -> Build (reference src/Manifold/Fish.cpp:13-95; SelfAttention::Build TGraph.cpp:94-165;
// FFN::Build EmbedVAE.cpp:427-480; tensor names NN2NAME QWen.cpp:61-145)
int Fish::Build() {
    std::string* hFishErr = &error;
    const MODEL_CARD& c = config;
    const int W = tp_world;
    if (W < 1 || tp_rank < 0 || tp_rank >= W) return KF_ERR_BAD_ARG;
    if (c.n_head % W || c.n_head_kv % W || c.n_ff % W || c.vocab % (16 * W) || (c.n_ff / W) % 128 || ((c.n_head / W) * c.head_dim) % 128) {
        error = "tensor-parallel degree must divide Head, KVHead, Ffn (in 128-wide groups) and vocab";
        return KF_ERR_BAD_ARG;
    }
    const int E = c.n_embd, hd = c.head_dim, nh = c.n_head / W, nkv = c.n_head_kv / W, ff = c.n_ff / W;
    const int QD = nh * hd, KD = nkv * hd;
    embed.hFish = this, embed.name = "model.embed_tokens";
    AllocTensor("model.embed_tokens.weight", c.vocab, E, 0, embed.w);
    cls.hFish = this, cls.name = "lm_head";
    cls.norm.hFish = this, cls.norm.rms_eps = c.norm_rms_eps;
    AllocTensor("model.norm.weight", 1, E, 1, cls.norm.w);
    cls.proj.hFish = this, cls.proj.nIn = E, cls.proj.nOut = c.vocab;
    if (c.tie_word_embeddings)
        cls.proj.w = embed.w;
    else
        AllocTensor("lm_head.weight", c.vocab, E, 2, cls.proj.w);
    for (int l = 0; l < c.n_layers; l++) {
        const std::string p = "model.layers." + std::to_string(l) + ".";
        const int b         = 16 + 16 * l;
        auto a = std::make_unique<SelfAttention>();
        a->hFish = this, a->name = p + "self_attn", a->layid = l + 1;
        a->n_head = nh, a->n_head_kv = nkv, a->head_dim = hd;
        a->norm.hFish = this, a->norm.rms_eps = c.norm_rms_eps;
        AllocTensor(p + "input_layernorm.weight", 1, E, b + 0, a->norm.w);
        SLP* slps[4]      = {&a->Q, &a->K, &a->V, &a->proj_cat};
        const char* nm[4] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight"};
        const int rr[4] = {QD, KD, KD, E}, cc[4] = {E, E, E, QD}, ids[4] = {b + 1, b + 2, b + 3, b + 6};
        for (int i = 0; i < 4; i++) {
            slps[i]->hFish = this, slps[i]->nIn = cc[i], slps[i]->nOut = rr[i];
            AllocTensor(p + nm[i], rr[i], cc[i], ids[i], slps[i]->w);
        }
        a->rope.hFish = this, a->rope.theta = c.rope_theta;
        if (c.isQKNormal) {
            AllocTensor(p + "self_attn.q_norm.weight", 1, hd, b + 4, a->rope.q_norm);
            AllocTensor(p + "self_attn.k_norm.weight", 1, hd, b + 5, a->rope.k_norm);
        }
        attn.push_back(std::move(a));
        auto m = std::make_unique<FFN>();
        m->hFish = this, m->name = p + "mlp", m->layid = l + 1, m->latent = ff;
        m->norm.hFish = this, m->norm.rms_eps = c.norm_rms_eps;
        AllocTensor(p + "post_attention_layernorm.weight", 1, E, b + 7, m->norm.w);
        SLP* fs[3]        = {&m->gate, &m->up, &m->down};
        const char* fn[3] = {"mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight"};
        const int fr[3] = {ff, ff, E}, fc[3] = {E, E, ff}, fi[3] = {b + 8, b + 9, b + 10};
        for (int i = 0; i < 3; i++) {
            fs[i]->hFish = this, fs[i]->nIn = fc[i], fs[i]->nOut = fr[i];
            AllocTensor(p + fn[i], fr[i], fc[i], fi[i], fs[i]->w);
        }
        ffn.push_back(std::move(m));
    }
    // KV cache (Fish::AllocBuffer, Fish.cpp:895-933) + activations
    cache.n_layer = c.n_layers, cache.max_batch = c.max_batch, cache.max_seq = c.max_seq_len, cache.kv_dim = KD;
    KF_TRY(kf_malloc(ctx, cache.bytes() / 2, &cache.key));
    KF_TRY(kf_malloc(ctx, cache.bytes() / 2, &cache.value));
    KF_TRY(kf_memset(ctx, cache.key, 0, cache.bytes() / 2));
    KF_TRY(kf_memset(ctx, cache.value, 0, cache.bytes() / 2));
    max_tokens = std::max(std::max(64, c.max_batch), c.max_prefill);
    logit_rows = std::max(64, c.max_batch);
    const size_t T = max_tokens;
    KF_TRY(kf_malloc(ctx, T * E * 2, &x));
    KF_TRY(kf_malloc(ctx, T * E * 2, &xb));
    KF_TRY(kf_malloc(ctx, T * QD * 2, &q));
    KF_TRY(kf_malloc(ctx, T * KD * 2, &k));
    KF_TRY(kf_malloc(ctx, T * KD * 2, &v));
    KF_TRY(kf_malloc(ctx, T * QD * 2, &att));
    KF_TRY(kf_malloc(ctx, T * ff * 2, &hb));
    KF_TRY(kf_malloc(ctx, (size_t)logit_rows * c.vocab * 2 * (W > 1 ? 2 : 1), &logits));
    if (W > 1) KF_TRY(kf_malloc(ctx, T * E * 4, (void**)&part_f32));
    KF_TRY(kf_malloc(ctx, T * 4, (void**)&d_tokens));
    KF_TRY(kf_malloc(ctx, T * 4, (void**)&d_pos));
    KF_TRY(kf_malloc(ctx, T * 4, (void**)&d_next));
    KF_TRY(kf_host_alloc(T * 4 * 3, (void**)&h_stage));
    KF_TRY(kf_host_alloc((size_t)logit_rows * c.vocab * 2, (void**)&h_logits));
    // one rope table shared by all layers
    void* table = nullptr;
    KF_TRY(kf_malloc(ctx, (size_t)c.max_seq_len * (hd / 2) * 8, &table));
    KF_TRY(kf_rope_table(ctx, table, c.max_seq_len, hd, c.rope_theta));
    for (auto& a : attn) a->rope.table = table;
    rope_table_shared = table;
    attn_hint = c.max_seq_len - 1;
    if (const char* e = getenv("KF_GQA_MIN_BATCH")) gqa_min_batch = atoi(e);  // tuning sweeps only
    return KF_OK;
}

// which window of the full (unsharded) tensor does this rank hold?
static void shard_window(const Fish& f, const std::string& name, int rows_l, int cols_l, int* rows_g, int* cols_g, int* row0, int* col0) {
    const int W = f.tp_world, r = f.tp_rank;
    *rows_g = rows_l, *cols_g = cols_l, *row0 = 0, *col0 = 0;
    auto has = [&](const char* s) { return name.find(s) != std::string::npos; };
    if (has("q_proj") || has("k_proj") || has("v_proj") || has("gate_proj") || has("up_proj"))
        *rows_g = rows_l * W, *row0 = r * rows_l;  // column-parallel: split output rows
    else if (has("o_proj") || has("down_proj"))
        *cols_g = cols_l * W, *col0 = r * cols_l;  // row-parallel: split input columns (whole 128-wide groups)
}

// Host-only shard plan (no device): full shape of a tensor by its HF name and the window rank `rank` of `world` holds.
// Megatron-style: Q/K/V/gate/up split by output rows (heads / ffn), O/down by input columns in whole 128-wide quant groups, the
// rest replicated.  Returns false for an unknown name or an indivisible configuration.
bool ShardPlan(const MODEL_CARD& c, const std::string& name, int rank, int world, int* rows_g, int* cols_g, int* rows_l, int* cols_l, int* row0,
               int* col0) {
    if (world < 1 || rank < 0 || rank >= world) return false;
    if (c.n_head % world || c.n_head_kv % world || c.n_ff % world || (c.n_ff / world) % 128 || ((c.n_head / world) * c.head_dim) % 128) return false;
    auto has = [&](const char* s) { return name.find(s) != std::string::npos; };
    const int E = c.n_embd, QD = c.q_dim(), KD = c.kv_dim(), F = c.n_ff;
    int R, C;
    if (has("q_proj")) R = QD, C = E;
    else if (has("k_proj") || has("v_proj")) R = KD, C = E;
    else if (has("o_proj")) R = E, C = QD;
    else if (has("gate_proj") || has("up_proj")) R = F, C = E;
    else if (has("down_proj")) R = E, C = F;
    else if (has("embed_tokens") || has("lm_head")) R = c.vocab, C = E;
    else if (has("q_norm") || has("k_norm")) R = 1, C = c.head_dim;
    else if (has("norm")) R = 1, C = E;
    else return false;
    *rows_g = R, *cols_g = C, *rows_l = R, *cols_l = C, *row0 = 0, *col0 = 0;
    if (has("q_proj") || has("k_proj") || has("v_proj") || has("gate_proj") || has("up_proj"))
        *rows_l = R / world, *row0 = rank * (R / world);
    else if (has("o_proj") || has("down_proj"))
        *cols_l = C / world, *col0 = rank * (C / world);
    return true;
}

// huTensor::InitParam random path (reference src/Device/CUDA/huTensor.cu:157-231) + LowBit_worker at load
int Fish::InitParamRandom() {
    std::string* hFishErr = &error;
    size_t most = 0;
    for (auto& kv : tensors) most = std::max(most, kv.second->size());
    void* scratch = nullptr;
    KF_TRY(kf_malloc(ctx, most * 2, &scratch));
    weight_bytes = 0;
    for (auto& kv : tensors) {
        hGTensor t      = kv.second;
        const int id    = tensor_ids[kv.first];
        const uint64_t seed = (uint64_t)config.seed * 1000003ull + (uint64_t)id;
        int rg, cg, r0, c0;
        shard_window(*this, kv.first, t->ne[0], t->ne[1], &rg, &cg, &r0, &c0);
        int rc;
        if (t->ne[0] == 1) {  // norm weights: FIX_1 (or 1 + norm_sigma*z for tests)
            rc = kf_fill_normal(ctx, scratch, t->size(), seed, config.norm_sigma, 1.0f);
            if (!rc) rc = t->SetBF16FromDevice(scratch);
        } else {
            rc = kf_fill_normal_2d(ctx, scratch, t->ne[0], t->ne[1], (size_t)cg, (size_t)r0, (size_t)c0, seed, config.init_sigma, 0.f);
            if (!rc) rc = t->hQuant ? t->hQuant->LowBit_worker(t, scratch, 0x100) : t->SetBF16FromDevice(scratch);
        }
        if (rc) {
            error = "InitParam(" + kv.first + ") -> " + kf_status_string(rc) + " : " + kf_last_error(ctx);
            kf_free(ctx, scratch);
            return rc;
        }
        weight_bytes += t->nByte();
    }
    KF_TRY(kf_ctx_sync(ctx));
    KF_TRY(kf_free(ctx, scratch));
    ResetGraphs();
    return KF_OK;
}
int Fish::SetTensor(const std::string& name, const void* host, int rows, int cols) {
    std::string* hFishErr = &error;
    auto it = tensors.find(name);
    if (it == tensors.end()) {
        error = "unknown tensor '" + name + "'";
        return KF_ERR_BAD_ARG;
    }
    hGTensor t = it->second;
    int rg, cg, r0, c0;
    shard_window(*this, name, t->ne[0], t->ne[1], &rg, &cg, &r0, &c0);
    if ((rows != rg || cols != cg) && !(t->ne[0] == 1 && (size_t)rows * cols == t->size())) {
        error = "tensor '" + name + "': expected full shape [" + std::to_string(rg) + "," + std::to_string(cg) + "]";
        return KF_ERR_BAD_ARG;
    }
    std::vector<uint16_t> shard(t->size());
    const uint16_t* src = (const uint16_t*)host;
    for (int r = 0; r < t->ne[0]; r++) memcpy(&shard[(size_t)r * t->ne[1]], src + (size_t)(r0 + r) * cg + c0, (size_t)t->ne[1] * 2);
    void* dev = nullptr;
    KF_TRY(kf_malloc(ctx, shard.size() * 2, &dev));
    KF_TRY(kf_h2d(ctx, dev, shard.data(), shard.size() * 2));
    int rc = (t->hQuant && t->ne[0] > 1) ? t->hQuant->LowBit_worker(t, dev, 0x100) : t->SetBF16FromDevice(dev);
    kf_ctx_sync(ctx);
    kf_free(ctx, dev);
    ResetGraphs();
    if (rc) error = "SetTensor(" + name + ") -> " + kf_status_string(rc) + " : " + kf_last_error(ctx);
    return rc;
}
// The window [out r0, r0 + OCl) x [in c0, c0 + ICl) of a full AWQ triple as ONE blob qweight || qzeros || scales (GTensor::AllocAWQ).  r0 and OCl
// are multiples of 8 (whole int32 words), c0 and ICl of 128 (whole groups).  Pure host code (the tensor-parallel shard plan of SURVEY 8e applied
// to the vendor layout); out_blob holds ICl * OCl / 2 + (ICl / 128) * (OCl / 8) * 4 + (ICl / 128) * OCl * 2 bytes.
void AwqShardWindow(const void* qweight, const void* qzeros, const void* scales, int IC, int OC, int r0, int OCl, int c0, int ICl, uint8_t* out_blob) {
    (void)IC;
    const size_t W8g = (size_t)OC / 8, W8l = (size_t)OCl / 8, w0 = (size_t)r0 / 8, g0 = (size_t)c0 / 128;
    uint32_t* qw = (uint32_t*)out_blob;
    uint32_t* qz = qw + (size_t)ICl * W8l;
    uint16_t* sc = (uint16_t*)(qz + (size_t)(ICl / 128) * W8l);
    const uint32_t* fqw = (const uint32_t*)qweight;
    const uint32_t* fqz = (const uint32_t*)qzeros;
    const uint16_t* fsc = (const uint16_t*)scales;
    for (int i = 0; i < ICl; i++) memcpy(qw + (size_t)i * W8l, fqw + (size_t)(c0 + i) * W8g + w0, W8l * 4);
    for (int g = 0; g < ICl / 128; g++) {
        memcpy(qz + (size_t)g * W8l, fqz + (g0 + g) * W8g + w0, W8l * 4);
        memcpy(sc + (size_t)g * OCl, fsc + (g0 + g) * (size_t)OC + r0, (size_t)OCl * 2);
    }
}
// gpt.awq_repack = 1: the same window re-laid-out at load into the library's own 4-bit storage -- PackedQ words over [out][in] rows (reference
// src/PackedQ.hpp:99-183: code i < 16 of a word in `high` at bit 60 - 4i, the rest in `low`; bytes 0-7 = low) + the bf16 gama array
// [R_SCALE][C_SCALE][ZERO][STEP] (GTensor.cpp:456-510) with step = bf16(scale), zero = bf16(zero_point * scale) per (out row, 128 input columns)
// -- so the decode GEMV / tcgen05 kernels of section 3.1 / 3.2 run on it at full speed.  The codes move unchanged (an AWQ group IS a PackedQ
// group: 128 consecutive input columns of one output row); the arithmetic becomes the RTN one, step * code - zero, with bf16 step / zero instead
// of (code - zero_point) * fp16 scale: off from CU_Q42X_awq by the bf16 rounding of the two products (tests state the bound).  Host code.
static inline uint16_t f32_to_bf16_host(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32_host(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            uint32_t m = man;
            do { e++, m <<= 1; } while (!(m & 0x400u));
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (m & 0x3ffu) << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | man << 13;
    } else {
        bits = sign | (exp + 112) << 23 | man << 13;
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
size_t AwqRepackBytes(int OCl, int ICl) { return (size_t)OCl * ICl / 2 + 2 * ((size_t)OCl + ICl + 2 * ((size_t)OCl * ICl / 128)); }
void AwqRepackWindow(const void* qweight, const void* qzeros, const void* scales, int IC, int OC, int r0, int OCl, int c0, int ICl, uint8_t* out_blob) {
    (void)IC;
    static const int kOrder[8] = {0, 4, 1, 5, 2, 6, 3, 7};  // AWQ_REVERSE_ORDER: element k of a word sits in nibble kOrder[k]
    const uint32_t* fqw = (const uint32_t*)qweight;
    const uint32_t* fqz = (const uint32_t*)qzeros;
    const uint16_t* fsc = (const uint16_t*)scales;
    const size_t W8g = (size_t)OC / 8, gpr = (size_t)ICl / 128, nG = (size_t)OCl * gpr;
    uint64_t* words = (uint64_t*)out_blob;  // {low, high} per 32 codes
    uint16_t* gama  = (uint16_t*)(out_blob + (size_t)OCl * ICl / 2);
    memset(gama, 0, 2 * ((size_t)OCl + ICl));  // R / C scales: unused (NO_NORMAL)
    uint16_t* gZero = gama + OCl + ICl;
    uint16_t* gStep = gZero + nG;
    for (int o = 0; o < OCl; o++) {
        const int oc = r0 + o, shift = 4 * kOrder[oc & 7];
        const size_t wcol = (size_t)oc / 8;
        for (size_t g = 0; g < gpr; g++) {
            const size_t gg = (size_t)c0 / 128 + g;
            const int z     = (int)((fqz[gg * W8g + wcol] >> shift) & 0xFu);
            const float sc  = f16_to_f32_host(fsc[gg * (size_t)OC + oc]);
            gStep[(size_t)o * gpr + g] = f32_to_bf16_host(sc);
            gZero[(size_t)o * gpr + g] = f32_to_bf16_host((float)z * sc);
        }
    }
    // codes: 32 input rows at a time, every vendor word read once (it holds one input row of 8 output columns); a few host threads share the blocks
    const int nBlock = ICl / 32, W8l = OCl / 8, w0 = r0 / 8;
    auto work = [&](int b0, int b1) {
        for (int w = b0; w < b1; w++)
            for (int wc = 0; wc < W8l; wc++) {
                uint64_t high[8] = {0, 0, 0, 0, 0, 0, 0, 0}, low[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 0; i < 32; i++) {
                    const uint32_t word = fqw[((size_t)c0 + (size_t)w * 32 + i) * W8g + w0 + wc];
                    for (int k = 0; k < 8; k++) {
                        const uint64_t q = (word >> (4 * kOrder[k])) & 0xFu;
                        if (i < 16)
                            high[k] |= q << (60 - 4 * i);
                        else
                            low[k] |= q << (60 - 4 * (i - 16));
                    }
                }
                for (int k = 0; k < 8; k++) {
                    uint64_t* dst = words + 2 * ((size_t)(wc * 8 + k) * nBlock + w);
                    dst[0] = low[k], dst[1] = high[k];
                }
            }
    };
    const int nThread = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, nBlock}));
    if (nThread == 1 || (size_t)OCl * ICl < (1u << 22)) {
        work(0, nBlock);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nThread; t++) pool.emplace_back(work, (int)((long long)nBlock * t / nThread), (int)((long long)nBlock * (t + 1) / nThread));
        for (auto& th : pool) th.join();
    }
}
// Vendor AWQ tensors (GeQuant::ExTensor + GTensor::LoadParam of .qweight / .qzeros / .scales, reference src/Tensor/GeQuant.cpp:144-200,
// src/Manifold/Serialize.cpp:145-230): the FULL (unsharded) arrays as the checkpoint stores them --
//   qweight int32 [IC][OC / 8], qzeros int32 [IC / 128][OC / 8], scales fp16 [IC / 128][OC]     (IC = in_features, OC = out_features)
// -- cut to this rank's window (Q/K/V/gate/up: a column range of every array; O/down: a row range, whole 128-row groups) and uploaded
// as one blob.  No arithmetic touches the codes: the device reads the vendor layout as it is (awq.cu).
int Fish::SetTensorAWQ(const std::string& name, const void* qweight, const void* qzeros, const void* scales, int IC, int OC) {
    std::string* hFishErr = &error;
    auto it = tensors.find(name);
    if (it == tensors.end()) {
        error = "unknown tensor '" + name + "'";
        return KF_ERR_BAD_ARG;
    }
    hGTensor t = it->second;
    if (!t->hQuant || t->hQuant->params.type != AWQ) {
        error = "tensor '" + name + "': the checkpoint holds it in the vendor AWQ layout but the quantizer card does not say \"quant_method\": \"awq\" for it";
        return KF_ERR_BAD_ARG;
    }
    int rg, cg, r0, c0;  // rows = OC, cols = IC
    shard_window(*this, name, t->ne[0], t->ne[1], &rg, &cg, &r0, &c0);
    if (OC != rg || IC != cg || !qweight || !qzeros || !scales) {
        error = "tensor '" + name + "': expected AWQ arrays of the full shape in_features " + std::to_string(cg) + ", out_features " + std::to_string(rg);
        return KF_ERR_BAD_ARG;
    }
    const int OCl = t->ne[0], ICl = t->ne[1];
    if (ICl % 128 || OCl % 32) {
        error = "tensor '" + name + "': the AWQ layout needs in_features in whole 128-row groups and out_features a multiple of 32 per rank";
        return KF_ERR_BAD_ARG;
    }
    int rc = config.awq_repack ? t->Alloc(typNUMBER::Q4, 128) : t->AllocAWQ();
    if (rc) {
        error = "tensor '" + name + "': " + kf_last_error(ctx);
        return rc;
    }
    std::vector<uint8_t> blob(t->nByte());
    if (config.awq_repack)  // the library's own 4-bit storage: the fast matmul kernels apply (see AwqRepackWindow)
        AwqRepackWindow(qweight, qzeros, scales, IC, OC, r0, OCl, c0, ICl, blob.data());
    else
        AwqShardWindow(qweight, qzeros, scales, IC, OC, r0, OCl, c0, ICl, blob.data());
    KF_TRY(kf_h2d(ctx, t->data, blob.data(), blob.size()));
    KF_TRY(kf_ctx_sync(ctx));
    t->qBias = 0;
    ResetGraphs();
    return KF_OK;
}
hGTensor Fish::GetTensor(const std::string& name) const {
    auto it = tensors.find(name);
    return it == tensors.end() ? nullptr : it->second;
}

// Fish::ForwardOnRLS (reference src/Manifold/gLLM.cpp:755-769): run every neuron's cuInfer in graph order
int Fish::ForwardOnRLS(int M, bool want_logits) {
    std::string* hFishErr = &error;
    // tensor-parallel decode of up to 8 tokens: the exchange after O / down is the epilogue of those matmuls (kf_tp.cuh): 5 launches
    // per block instead of 7
    tp_fuse = false;
    if (tp_world > 1) {
        KF_TRY(kf_tp_begin(ctx));
        tp_fuse = 2 * (int)attn.size() <= 256 && kf_exchange_fused_ready(ctx, M, config.n_embd) == 1;
        auto own_matmul = [](const hGTensor& t) { return t->type == typNUMBER::Q4_NF || t->type == typNUMBER::Q4_AWQ; };
        for (size_t l = 0; l < attn.size() && tp_fuse; l++)  // NormalFloat4 / AWQ weights have their own matmul (nf4.cu, awq.cu): stand-alone exchange
            tp_fuse = !own_matmul(attn[l]->proj_cat.w) && !own_matmul(ffn[l]->down.w);
    }
    KF_TRY(embed.cuInfer(x, M));
    for (size_t l = 0; l < attn.size(); l++) {
        KF_TRY(attn[l]->cuInfer(x, M));
        KF_TRY(ffn[l]->cuInfer(x, M));
    }
    if (want_logits) {
        if (last_only)  // prefill: only the last token of the panel feeds the sampler
            KF_TRY(cls.cuInfer_1(logits, (uint16_t*)x + (size_t)(M - 1) * config.n_embd, 1));
        else
            KF_TRY(cls.cuInfer_1(logits, x, M));
    }
    return KF_OK;
}

int Fish::PickNext(int rows) {
    if (samp_temperature == 0.f || samp_top_k == 1) return kf_argmax(ctx, d_next, logits, rows, config.vocab);
    return kf_sample(ctx, d_next, logits, rows, config.vocab, samp_temperature, samp_top_k, samp_top_p, d_rng, samp_selection);
}
int Fish::SetSampler(float temperature, int top_k, float top_p, uint64_t seed, int selection) {
    std::string* hFishErr = &error;
    if (temperature < 0.f || top_p <= 0.f || (selection != 0 && selection != 1)) {
        error = "sampler: temperature >= 0, top_p > 0, selection 0 / 1";
        return KF_ERR_BAD_ARG;
    }
    ResetGraphs();  // the captured steps carry the sampler's launch
    samp_temperature = temperature, samp_top_k = top_k, samp_top_p = top_p, samp_selection = selection;
    const size_t rows = (size_t)std::max(logit_rows, max_tokens);
    if (!d_rng) KF_TRY(kf_malloc(ctx, rows * 8, (void**)&d_rng));
    std::vector<uint64_t> st(rows, seed);
    KF_TRY(kf_h2d(ctx, d_rng, st.data(), rows * 8));
    KF_TRY(kf_ctx_sync(ctx));
    return KF_OK;
}

// graph key: M<<12 | ctx bucket<<7 | want_logits | seq_mode<<1 | feedback<<2 | argmax<<3 | last_only<<5 | consecutive<<6
// contexts are bucketed by powers of two (>= 512): the attention launch geometry (slices, kernel choice) follows the bucket
int Fish::CtxBucket() const {
    int b = 9;
    while ((1 << b) < staged_pos_max + 1 && (1 << b) < config.max_seq_len) b++;
    return b;
}
int Fish::UseGraph(int M, bool want_logits) {
    attn_hint = std::min(config.max_seq_len, 1 << CtxBucket()) - 1;
    return (M << 12) | (CtxBucket() << 7) | (int)want_logits | (seq_mode << 1) | ((int)last_only << 5) | ((int)panel_consecutive << 6);
}

int Fish::Forward(const int32_t* tokens, const int32_t* pos, int M, int mode, uint16_t* logits_out, int32_t* next_out) {
    std::string* hFishErr = &error;
    KF_TRY(kf_ctx_make_current(ctx));
    if (!tokens || !pos || M < 1 || M > max_tokens) {
        error = "Forward: 1 <= M <= " + std::to_string(max_tokens);
        return KF_ERR_BAD_ARG;
    }
    if (mode == 1 && M > config.max_batch) {
        error = "Forward: batched decode needs gpt.max_batch >= M";
        return KF_ERR_BAD_ARG;
    }
    for (int m = 0; m < M; m++) {
        if (tokens[m] < 0 || tokens[m] >= config.vocab || pos[m] < 0 || pos[m] >= config.max_seq_len) {
            error = "Forward: token id or position out of range";
            return KF_ERR_BAD_ARG;
        }
        h_stage[m] = tokens[m], h_stage[max_tokens + m] = pos[m];
    }
    if (!all_resident) {  // a model whose weights come tensor by tensor (SetTensor / SetTensorAWQ / a checkpoint) must have all of them
        for (auto& kv : tensors)
            if (!kv.second->data) {
                error = "Forward: tensor '" + kv.first + "' has no data yet (init_random, set_tensor or a checkpoint must provide every tensor)";
                return KF_ERR_BAD_ARG;
            }
        all_resident = true;
    }
    SyncGraphGeneration();
    staged_pos_max = *std::max_element(pos, pos + M);
    staged_M = M;
    seq_mode = mode == 1 ? 1 : 0;
    last_only = mode == 2;
    panel_consecutive = seq_mode == 0;
    for (int m = 1; m < M && panel_consecutive; m++) panel_consecutive = pos[m] == pos[0] + m;
    const int R = last_only ? 1 : M;  // rows of logits / argmax produced
    if ((logits_out || next_out) && R > logit_rows) {
        error = "Forward: logits of more than " + std::to_string(logit_rows) + " tokens requested; use seq_mode 2 (last token only) for long panels";
        return KF_ERR_BAD_ARG;
    }
    KF_TRY(kf_h2d(ctx, d_tokens, h_stage, (size_t)M * 4));
    KF_TRY(kf_h2d(ctx, d_pos, h_stage + max_tokens, (size_t)M * 4));
    const bool want_logits = logits_out || next_out;
    const int key          = UseGraph(M, want_logits) | ((next_out ? 1 : 0) << 3);
    auto it                = graphs.find(key);
    if (use_graphs && it == graphs.end() && warm.count(key)) {  // second call with this signature: capture it
        KF_TRY(kf_graph_begin(ctx));
        int rc = ForwardOnRLS(M, want_logits);
        if (!rc && next_out) rc = PickNext(R);
        kf_graph* g = nullptr;
        int rc2     = kf_graph_end(ctx, &g);
        if (rc || rc2) {
            error = std::string("graph capture failed: ") + kf_last_error(ctx);
            return rc ? rc : rc2;
        }
        graphs[key] = g;
        it          = graphs.find(key);
    }
    if (use_graphs && it != graphs.end()) {
        KF_TRY(kf_graph_launch(ctx, it->second));
    } else {
        KF_TRY(ForwardOnRLS(M, want_logits));
        if (next_out) KF_TRY(PickNext(R));
        warm.insert(key);
    }
    if (logits_out) KF_TRY(kf_d2h(ctx, h_logits, logits, (size_t)R * config.vocab * 2));
    if (next_out) KF_TRY(kf_d2h(ctx, h_stage + 2 * max_tokens, d_next, (size_t)R * 4));
    KF_TRY(kf_ctx_sync(ctx));
    if (logits_out) memcpy(logits_out, h_logits, (size_t)R * config.vocab * 2);
    if (next_out) memcpy(next_out, h_stage + 2 * max_tokens, (size_t)R * 4);
    if (last_only && next_out) {  // leave the model ready for DecodeLoop(n, 1): feed the sampled token at the next position
        h_stage[0] = next_out[0], h_stage[max_tokens] = pos[M - 1] + 1;
        KF_TRY(kf_h2d(ctx, d_tokens, h_stage, 4));
        KF_TRY(kf_h2d(ctx, d_pos, h_stage + max_tokens, 4));
        KF_TRY(kf_ctx_sync(ctx));
        staged_pos_max = pos[M - 1] + 1;
        staged_M = 1;
    }
    return KF_OK;
}

// The generation loop of Fish::Chat (reference src/Manifold/GoPT.cpp:1111-1235): prefill the prompt, then sample -> stop on eos or a full
// context window -> feed the token back.  The reference prefills token by token (one Evaluate per prompt token, :1140-1147); here the prompt
// goes through in panels of up to max_tokens (seq_mode 2: only the last token's logits are formed) and every generated token is one
// Forward() with a 4-byte result.  The token that equals eos_id is not emitted, as the reference does not print it (:1171).
int Fish::Generate(const int32_t* prompt, int n_prompt, int pos0, int max_new, int eos_id, int32_t* out, int* n_out, int* stop_reason) {
    if (n_out) *n_out = 0;
    if (stop_reason) *stop_reason = 0;
    if (!prompt || n_prompt < 1 || pos0 < 0 || max_new < 0 || (max_new > 0 && !out) || !n_out) {
        error = "Generate: needs a prompt of at least one token, pos0 >= 0, max_new_tokens >= 0 and an output buffer";
        return KF_ERR_BAD_ARG;
    }
    if ((long long)pos0 + n_prompt > config.max_seq_len) {
        error = "Generate: the prompt does not fit the context window (gpt.max_seq_len = " + std::to_string(config.max_seq_len) + ")";
        return KF_ERR_BAD_ARG;
    }
    std::vector<int32_t> posv((size_t)max_tokens);
    int32_t next = -1;
    for (int off = 0; off < n_prompt; off += max_tokens) {
        const int m = std::min(max_tokens, n_prompt - off);
        for (int i = 0; i < m; i++) posv[i] = pos0 + off + i;
        const bool last = off + m == n_prompt;
        const int rc    = Forward(prompt + off, posv.data(), m, 2, nullptr, last ? &next : nullptr);
        if (rc) return rc;
    }
    int32_t pos = pos0 + n_prompt;  // where the next token will sit
    int n       = 0;
    for (;;) {
        if (next == eos_id) {
            if (stop_reason) *stop_reason = 1;
            break;
        }
        if (n >= max_new) {
            if (stop_reason) *stop_reason = 2;
            break;
        }
        out[n++] = next;
        if (n >= max_new) {
            if (stop_reason) *stop_reason = 2;
            break;
        }
        if (pos >= config.max_seq_len) {  // "context window full!" (GoPT.cpp:1175)
            if (stop_reason) *stop_reason = 3;
            break;
        }
        int32_t fed = next;
        const int rc = Forward(&fed, &pos, 1, 0, nullptr, &next);
        if (rc) return rc;
        pos++;
    }
    *n_out = n;
    return KF_OK;
}

// Device-resident greedy decoding: each step = one CUDA-graph replay of [forward, argmax, token feedback, pos++].  The tokens and
// positions staged by the last Forward() call are the starting state.
int Fish::DecodeLoop(int n_steps, int M) {
    std::string* hFishErr = &error;
    KF_TRY(kf_ctx_make_current(ctx));
    if (M < 1 || M > max_tokens || n_steps < 0) {
        error = "DecodeLoop: 1 <= M <= " + std::to_string(max_tokens) + ", n_steps >= 0";
        return KF_ERR_BAD_ARG;
    }
    if (staged_M < M) {  // the loop continues from what the last Forward() staged: there must be M (token, position) pairs on the device
        error = "DecodeLoop: call Forward() with at least " + std::to_string(M) + " token(s) first (it stages the starting tokens / positions)";
        return KF_ERR_BAD_ARG;
    }
    if (M > 1 && (M > config.max_batch || M > logit_rows)) {  // M independent sequences: each needs its own KV region and logits row
        error = "DecodeLoop: batched decode needs gpt.max_batch >= M (and at most " + std::to_string(logit_rows) + " sequences)";
        return KF_ERR_BAD_ARG;
    }
    if (staged_pos_max + n_steps >= config.max_seq_len) {
        error = "DecodeLoop: would run past gpt.max_seq_len";
        return KF_ERR_BAD_ARG;
    }
    SyncGraphGeneration();
    staged_pos_max += n_steps;
    seq_mode = M > 1 ? 1 : seq_mode, last_only = false, panel_consecutive = false;
    const int key = UseGraph(M, true) | (1 << 2);
    auto it       = graphs.find(key);
    auto body     = [&]() -> int {
        int rc = ForwardOnRLS(M, true);
        if (!rc) rc = PickNext(M);
        if (!rc) rc = kf_d2d(ctx, d_tokens, d_next, (size_t)M * 4);
        if (!rc) rc = kf_advance_pos(ctx, d_pos, M);
        return rc;
    };
    int done = 0;
    if (it == graphs.end()) {
        if (n_steps == 0) return KF_OK;
        KF_TRY(body());  // eager first step sizes every workspace
        done = 1;
        SyncGraphGeneration();  // ... which may have re-allocated scratch that older graphs point to
        if (use_graphs) {
            KF_TRY(kf_graph_begin(ctx));
            int rc      = body();
            kf_graph* g = nullptr;
            int rc2     = kf_graph_end(ctx, &g);
            if (rc || rc2) {
                error = std::string("graph capture failed: ") + kf_last_error(ctx);
                return rc ? rc : rc2;
            }
            graphs[key] = g;
            it          = graphs.find(key);
        }
    }
    for (; done < n_steps; done++) {
        if (use_graphs && it != graphs.end())
            KF_TRY(kf_graph_launch(ctx, it->second));
        else
            KF_TRY(body());
    }
    return KF_OK;
}

// ---- serialisation of the resident tensors: one record per tensor, payload = the device blob `data || gama` byte for byte (what the
// reference writes per tensor in SerialGamaData, src/Device/CUDA/huTensor.cu:413-458, inside its fish.kun container).  Loading skips the
// quantiser entirely.  Layout: "KFB1" u32 count, then per tensor: u32 name_len, name, i32 type, rows, cols, group, qbias, u64 szData,
// u64 szGama, payload.  Tensor-parallel ranks save / load their own shard files.
int Fish::SaveBlobs(const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) {
        error = "cannot open '" + path + "' for writing";
        return KF_ERR_BAD_ARG;
    }
    std::string* hFishErr = &error;
    (void)hFishErr;
    const uint32_t magic = 0x3142464bu, count = (uint32_t)tensors.size();  // "KFB1"
    fwrite(&magic, 4, 1, f), fwrite(&count, 4, 1, f);
    std::vector<uint8_t> host;
    int rc = KF_OK;
    for (auto& kv : tensors) {
        const hGTensor& t = kv.second;
        const uint32_t nl = (uint32_t)kv.first.size();
        const int32_t meta[5] = {(int32_t)t->type, t->ne[0], t->ne[1], (t->hQuant && t->ne[0] > 1) ? t->hQuant->params.T_group : 0, t->qBias};
        const uint64_t sz[2] = {t->szData, t->szGama};
        host.resize(t->nByte());
        rc = kf_d2h(ctx, host.data(), t->data, t->nByte());
        if (!rc) rc = kf_ctx_sync(ctx);
        if (rc) break;
        fwrite(&nl, 4, 1, f), fwrite(kv.first.data(), 1, nl, f), fwrite(meta, 4, 5, f), fwrite(sz, 8, 2, f);
        if (fwrite(host.data(), 1, host.size(), f) != host.size()) {
            error = "short write to '" + path + "'";
            rc    = KF_ERR_BAD_ARG;
            break;
        }
    }
    fclose(f);
    if (rc && error.empty()) error = std::string("SaveBlobs: ") + kf_last_error(ctx);
    return rc;
}
int Fish::LoadBlobs(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        error = "cannot open '" + path + "'";
        return KF_ERR_BAD_ARG;
    }
    auto fail = [&](const std::string& why) {
        fclose(f);
        error = "LoadBlobs('" + path + "'): " + why;
        return KF_ERR_BAD_ARG;
    };
    uint32_t magic = 0, count = 0;
    if (fread(&magic, 4, 1, f) != 1 || fread(&count, 4, 1, f) != 1 || magic != 0x3142464bu) return fail("not a KFB1 file");
    if (count != tensors.size()) return fail("tensor count differs from the model built from this config");
    std::vector<uint8_t> host;
    for (uint32_t i = 0; i < count; i++) {
        uint32_t nl = 0;
        if (fread(&nl, 4, 1, f) != 1 || nl > 4096) return fail("corrupt record header");
        std::string name(nl, '\0');
        int32_t meta[5];
        uint64_t sz[2];
        if (fread(&name[0], 1, nl, f) != nl || fread(meta, 4, 5, f) != 5 || fread(sz, 8, 2, f) != 2) return fail("truncated record");
        auto it = tensors.find(name);
        if (it == tensors.end()) return fail("unknown tensor '" + name + "'");
        const hGTensor& t = it->second;
        // what this model's config selects for the tensor: quantised (bits, group of its card) or plain bf16 (norms, or a quantiser
        // that leaves vectors alone).  A model that has not been initialised yet gets its device blob allocated here.
        const typNUMBER tp  = (typNUMBER)meta[0];
        const bool quantised = t->hQuant && t->ne[0] > 1;
        const int group      = quantised ? t->hQuant->params.T_group : 0;
        const bool type_ok   = quantised ? ((int)BitPE(tp) == t->hQuant->bits && tp != typNUMBER::BF16) || (t->hQuant->bits == 16 && tp == typNUMBER::BF16)
                                         : tp == typNUMBER::BF16;
        if (!type_ok || meta[1] != t->ne[0] || meta[2] != t->ne[1] || meta[3] != group)
            return fail("tensor '" + name + "' was saved with another shape / storage type / group than this config selects");
        if (!t->data || t->type != tp || t->szData != sz[0] || t->szGama != sz[1]) {
            int rc = t->Alloc(tp, group);
            if (rc) {
                fclose(f);
                error = std::string("LoadBlobs: ") + kf_last_error(ctx);
                return rc;
            }
        }
        if (sz[0] != t->szData || sz[1] != t->szGama) return fail("tensor '" + name + "': byte sizes do not match its shape and type");
        host.resize(t->nByte());
        if (fread(host.data(), 1, host.size(), f) != host.size()) return fail("truncated payload of '" + name + "'");
        t->qBias = meta[4];
        int rc   = kf_h2d(ctx, t->data, host.data(), host.size());
        if (!rc) rc = kf_ctx_sync(ctx);
        if (rc) {
            fclose(f);
            error = std::string("LoadBlobs: ") + kf_last_error(ctx);
            return rc;
        }
    }
    fclose(f);
    ResetGraphs();
    return KF_OK;
}

// ---- the reference's own container, fish.kun (CKP_KOIFISH; csrc/Tensor/KunFile.cpp has the format): every resident tensor as
// {"dtype": K_FLOATS name, "shape", "data_offsets", "loAB", "szGama", "szData"} + payload data || gama, and the model config as the msgpack
// "__koifish__config__" entry.  What Fish::SAFETENSOR_Serialize writes / SAFETENSOR2Gensors + GTensor::LoadParam + Serial_Quant_MMAP read
// (reference src/Manifold/Serialize.cpp:145-230, 770-1010; src/Device/CUDA/huTensor.cu:487-588) for the inference tensors.
static const char* kunDtype(typNUMBER t) {  // K_FLOATS, src/g_float.hpp:127-151
    switch (t) {
        case typNUMBER::BF16: return "BF16(E8)";
        case typNUMBER::F8E5M2: return "F8E5M2";
        case typNUMBER::Q4:
        case typNUMBER::Q4_NF: return "Q<4>";  // NormalFloat4 is typNUMBER::Q4 under QUANT_MODE::RTNf: the quant card tells them apart
        case typNUMBER::Q2: return "Q<2>";
        case typNUMBER::T_SIGN: return "TERNARY";
        case typNUMBER::T_BINARY: return "BINARY";
        case typNUMBER::Q4_AWQ: return nullptr;
    }
    return nullptr;
}
int Fish::SaveKun(const std::string& path, const std::string& config_json) {
    std::vector<std::vector<uint8_t>> host(tensors.size());
    std::vector<KunTensorOut> outs;
    size_t i = 0;
    for (auto& kv : tensors) {
        const hGTensor& t = kv.second;
        const char* dt    = kunDtype(t->type);
        if (!t->data || !dt) {
            error = !t->data ? "SaveKun: tensor '" + kv.first + "' has no data"
                             : "SaveKun: '" + kv.first + "' is in the vendor AWQ layout -- it already has a checkpoint format of its own (.qweight / .qzeros / .scales)";
            return KF_ERR_BAD_ARG;
        }
        host[i].resize(t->nByte());
        int rc = kf_d2h(ctx, host[i].data(), t->data, t->nByte());
        if (!rc) rc = kf_ctx_sync(ctx);
        if (rc) {
            error = std::string("SaveKun: ") + kf_last_error(ctx);
            return rc;
        }
        KunTensorOut o;
        o.name = kv.first, o.dtype = dt;
        if (t->ne[0] == 1)
            o.shape[0] = t->ne[1], o.shape[1] = 0;  // norm weights are vectors
        else
            o.shape[0] = t->ne[0], o.shape[1] = t->ne[1];
        o.szData = t->szData, o.szGama = t->szGama, o.blob = host[i].data();
        outs.push_back(o);
        i++;
    }
    std::string err;
    if (kun_write(path, config_json, outs, &err) != 0) {
        error = "SaveKun('" + path + "'): " + err;
        return KF_ERR_BAD_ARG;
    }
    return KF_OK;
}
int Fish::LoadKun(const std::string& path, int* n_loaded, int* n_skipped) {
    KunFile file;
    std::string err;
    if (kun_parse(path, &file, &err) != 0) {
        error = "LoadKun: " + err;
        return KF_ERR_BAD_ARG;
    }
    auto fail = [&](const std::string& why) {
        error = "LoadKun('" + path + "'): " + why;
        return KF_ERR_BAD_ARG;
    };
    int loaded = 0, skipped = 0;
    std::vector<uint8_t> host;
    for (const KunEntry& e : file.entries) {
        auto it = tensors.find(e.name);
        if (it == tensors.end()) {  // the reference treats an unknown key as an error (Serialize.cpp:800-804); a tied "model.out.weight" or
            skipped++;              // rotary tables of another writer are harmless here, so they are counted instead
            continue;
        }
        const hGTensor& t   = it->second;
        const bool quantised = t->hQuant && t->ne[0] > 1;
        typNUMBER want = quantised ? t->hQuant->params.tpQuant() : typNUMBER::BF16;
        if (want == typNUMBER::Q4_AWQ && config.awq_repack) want = typNUMBER::Q4;  // repacked at load: stored (and saved) as the library's own Q4
        const char* dt = kunDtype(want);
        // accept the HF spelling for plain tensors ("BF16"), otherwise the K_FLOATS name this model's config selects for the tensor
        if (!dt || !(e.dtype == dt || (want == typNUMBER::BF16 && e.dtype == "BF16")))
            return fail("tensor '" + e.name + "' is stored as " + e.dtype + ", this config selects " + (dt ? dt : "the AWQ layout"));
        uint64_t numel = 1;
        for (int64_t d : e.shape) numel *= (uint64_t)d;
        const bool shape_ok = t->ne[0] == 1 ? numel == t->size()
                                            : (e.shape.size() == 2 && e.shape[0] == t->ne[0] && e.shape[1] == t->ne[1]);
        if (!shape_ok) return fail("tensor '" + e.name + "' has another shape than the model built from this config (tensor-parallel ranks use one file per rank)");
        if (!t->data || t->type != want) {
            const int rc = t->Alloc(want, quantised ? t->hQuant->params.T_group : 0);
            if (rc) {
                error = std::string("LoadKun: ") + kf_last_error(ctx);
                return rc;
            }
        }
        if (e.szData != t->szData || e.szGama != t->szGama)
            return fail("tensor '" + e.name + "': szData / szGama (" + std::to_string(e.szData) + " / " + std::to_string(e.szGama) + ") differ from this config's (" +
                        std::to_string(t->szData) + " / " + std::to_string(t->szGama) + "): another group size?");
        host.resize(t->nByte());
        if (kun_read(file, e, host.data(), &err) != 0) return fail(err);
        t->qBias = quantised ? t->hQuant->qBias : 0;
        int rc = kf_h2d(ctx, t->data, host.data(), host.size());
        if (!rc) rc = kf_ctx_sync(ctx);
        if (rc) {
            error = std::string("LoadKun: ") + kf_last_error(ctx);
            return rc;
        }
        loaded++;
    }
    if (n_loaded) *n_loaded = loaded;
    if (n_skipped) *n_skipped = skipped;
    ResetGraphs();
    return KF_OK;
}

}  // namespace koifish
