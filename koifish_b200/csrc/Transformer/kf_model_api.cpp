// kf_model_api.cpp -- extern "C" surface of the host runtime (include/kf_model.h).  Exceptions never cross the boundary.
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../Tensor/KunFile.hpp"
#include "../Tensor/Safetensors.hpp"
#include "QWen3.hpp"
#include "kf_model.h"

using namespace koifish;

struct kf_model {
    std::unique_ptr<Fish> fish;
    std::vector<std::string> names;
    std::string config_text;  // the JSON the model was created from (written into fish.kun files)
};

static char* dup_cstr(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

extern "C" int kf_model_create(kf_ctx* ctx, const char* config_json, int tp_rank, int tp_world, kf_model** out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!ctx || !config_json || !out) return KF_ERR_BAD_ARG;
    *out = nullptr;
    try {
        JSON j          = JSON::parse(config_json);
        MODEL_CARD card = MODEL_CARD::FromJSON(j);
        auto m          = std::make_unique<kf_model>();
        m->fish         = std::make_unique<Fish>(ctx, card, tp_rank, tp_world);
        int rc          = m->fish->Build();
        if (rc) {
            if (err_out) *err_out = dup_cstr(m->fish->error + " : " + kf_last_error(ctx));
            return rc;
        }
        for (auto& kv : m->fish->tensors) m->names.push_back(kv.first);
        m->config_text = json_dump(j);
        *out = m.release();
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_destroy(kf_model* m) {
    if (!m) return KF_ERR_BAD_ARG;
    delete m;
    return KF_OK;
}
extern "C" const char* kf_model_error(kf_model* m) { return m ? m->fish->error.c_str() : ""; }
extern "C" void kf_string_free(char* s) { free(s); }

extern "C" int kf_model_info_get(kf_model* m, kf_model_info* o) {
    if (!m || !o) return KF_ERR_BAD_ARG;
    Fish& f             = *m->fish;
    const MODEL_CARD& c = f.config;
    memset(o, 0, sizeof(*o));
    o->n_layers = c.n_layers, o->n_embd = c.n_embd, o->n_ff = c.n_ff, o->n_head = c.n_head, o->n_head_kv = c.n_head_kv;
    o->head_dim = c.head_dim, o->vocab = c.vocab, o->max_seq_len = c.max_seq_len, o->max_batch = c.max_batch, o->max_tokens = f.max_tokens;
    o->tp_rank = f.tp_rank, o->tp_world = f.tp_world, o->tie_word_embeddings = c.tie_word_embeddings;
    o->rope_theta = c.rope_theta, o->norm_rms_eps = c.norm_rms_eps;
    if (f.weights_dirty) {  // what is resident now (random init, set tensor by tensor, or from a checkpoint)
        f.weight_bytes = 0;
        for (auto& kv : f.tensors)
            if (kv.second->data) f.weight_bytes += kv.second->nByte();
        f.weights_dirty = false;
    }
    o->weight_bytes = f.weight_bytes, o->kv_bytes = f.cache.bytes();
    if (!f.attn.empty()) {
        auto nb = [](const hGTensor& t) { return t ? (uint64_t)t->nByte() : 0ull; };
        SelfAttention& a = *f.attn[0];
        FFN& n           = *f.ffn[0];
        o->block_weight_bytes_per_layer = nb(a.norm.w) + nb(a.Q.w) + nb(a.K.w) + nb(a.V.w) + nb(a.proj_cat.w) + nb(a.rope.q_norm) + nb(a.rope.k_norm) +
                                          nb(n.norm.w) + nb(n.gate.w) + nb(n.up.w) + nb(n.down.w);
    }
    if (f.cls.proj.w) o->head_weight_bytes = (uint64_t)f.cls.proj.w->nByte() / (uint64_t)f.tp_world;
    return KF_OK;
}
extern "C" int kf_model_init_random(kf_model* m) {
    if (!m) return KF_ERR_BAD_ARG;
    try {
        return m->fish->InitParamRandom();
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_set_tensor(kf_model* m, const char* name, const void* host, int rows, int cols) {
    if (!m || !name || !host) return KF_ERR_BAD_ARG;
    try {
        return m->fish->SetTensor(name, host, rows, cols);
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_set_tensor_awq(kf_model* m, const char* name, const void* qweight, const void* qzeros, const void* scales, int in_features,
                                       int out_features) {
    if (!m || !name) return KF_ERR_BAD_ARG;
    try {
        return m->fish->SetTensorAWQ(name, qweight, qzeros, scales, in_features, out_features);
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_tensor_desc(kf_model* m, const char* name, kf_tensor_desc* out) {
    if (!m || !name || !out) return KF_ERR_BAD_ARG;
    hGTensor t = m->fish->GetTensor(name);
    if (!t || !t->data) return KF_ERR_BAD_ARG;
    *out = t->Desc();
    return KF_OK;
}
extern "C" int kf_model_tensor_count(kf_model* m) { return m ? (int)m->names.size() : 0; }
extern "C" const char* kf_model_tensor_name(kf_model* m, int i) { return (m && i >= 0 && i < (int)m->names.size()) ? m->names[i].c_str() : nullptr; }
extern "C" void* kf_model_kcache(kf_model* m, int layer) {
    return (m && layer >= 0 && layer < m->fish->cache.n_layer) ? m->fish->cache.Get(KVCache::KV_KEY, layer) : nullptr;
}
extern "C" void* kf_model_vcache(kf_model* m, int layer) {
    return (m && layer >= 0 && layer < m->fish->cache.n_layer) ? m->fish->cache.Get(KVCache::KV_VAL, layer) : nullptr;
}
extern "C" int kf_model_forward(kf_model* m, const int32_t* tokens, const int32_t* pos, int M, int seq_mode, void* logits, int32_t* next) {
    if (!m) return KF_ERR_BAD_ARG;
    try {
        return m->fish->Forward(tokens, pos, M, seq_mode, (uint16_t*)logits, next);
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_decode_loop(kf_model* m, int n_steps, int M) {
    if (!m) return KF_ERR_BAD_ARG;
    try {
        return m->fish->DecodeLoop(n_steps, M);
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_read_state(kf_model* m, int32_t* tokens, int32_t* pos, int M) {
    if (!m || M < 1 || M > m->fish->max_tokens) return KF_ERR_BAD_ARG;
    Fish& f = *m->fish;
    int rc  = KF_OK;
    if (tokens) rc = kf_d2h(f.ctx, f.h_stage, f.d_tokens, (size_t)M * 4);
    if (!rc && pos) rc = kf_d2h(f.ctx, f.h_stage + f.max_tokens, f.d_pos, (size_t)M * 4);
    if (!rc) rc = kf_ctx_sync(f.ctx);
    if (rc) return rc;
    if (tokens) memcpy(tokens, f.h_stage, (size_t)M * 4);
    if (pos) memcpy(pos, f.h_stage + f.max_tokens, (size_t)M * 4);
    return KF_OK;
}
extern "C" int kf_config_dims(const char* config_json, kf_model_info* o, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !o) return KF_ERR_BAD_ARG;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        memset(o, 0, sizeof(*o));
        o->n_layers = c.n_layers, o->n_embd = c.n_embd, o->n_ff = c.n_ff, o->n_head = c.n_head, o->n_head_kv = c.n_head_kv;
        o->head_dim = c.head_dim, o->vocab = c.vocab, o->max_seq_len = c.max_seq_len, o->max_batch = c.max_batch;
        o->tp_world = 1, o->tie_word_embeddings = c.tie_word_embeddings, o->rope_theta = c.rope_theta, o->norm_rms_eps = c.norm_rms_eps;
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_BAD_ARG;
    }
}
// host only: the "quantizer" block the config resolves to, as JSON text -- for an HF config with "quantization_config" this is what
// QUANT_CARD::Vendor2JSONx builds (reference src/Utils/CLI_params.cpp:240-262); "" when the config quantises nothing
extern "C" int kf_config_quantizer_json(const char* config_json, char** json_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !json_out) return KF_ERR_BAD_ARG;
    *json_out = nullptr;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        *json_out    = dup_cstr(c.jQuant.is_null() ? std::string() : json_dump(c.jQuant));
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_UNSUPPORTED;
    }
}
// host only: the quantizer card Init4Neuron fills for a tensor name, field by field -- out[8] = {selected, mode (0 none, 1 RTN, 2 AWQ, 3 RTNf,
// 5 F8Ex: the reference's QUANT_MODE numbering, src/CLI_params.hpp:479-492), default_bits, T_group, yyang, isSymmetric, isZeroPoint, isVendorQuant}
extern "C" int kf_config_quant_card(const char* config_json, const char* tensor_name, int* out, float* errq_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !tensor_name || !out) return KF_ERR_BAD_ARG;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        QUANT_CARD card;
        const bool sel = card.Init4Neuron(tensor_name, c.jQuant);
        const int mode = card.type == RTN ? 1 : card.type == AWQ ? 2 : card.type == RTNf ? 3 : card.type == F8Ex ? 5 : 0;
        out[0] = sel, out[1] = mode, out[2] = card.default_bits, out[3] = card.T_group, out[4] = (int)card.yyang, out[5] = card.isSymmetric,
        out[6] = card.isZeroPoint, out[7] = card.isVendorQuant;
        if (errq_out) *errq_out = card.T_errQ;
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_UNSUPPORTED;
    }
}
extern "C" int kf_config_quant_of(const char* config_json, const char* tensor_name, int* type_out, int* group_out, int* mode_out,
                                  int* qbias_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !tensor_name) return KF_ERR_BAD_ARG;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        hQUANT q     = GeQuant::MakeInstance(tensor_name, c.jQuant);
        if (type_out) *type_out = q ? kfType(q->params.tpQuant()) : KF_T_BF16;
        if (group_out) *group_out = q ? q->params.T_group : 0;
        if (mode_out) *mode_out = q ? q->params.kfMode() : 0;
        if (qbias_out) *qbias_out = q ? q->qBias : 0;
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_UNSUPPORTED;
    }
}
// Host only: rank `rank` of `world`'s blob (qweight || qzeros || scales, what kf_model_set_tensor_awq uploads) of the AWQ linear `tensor_name`
// given its FULL arrays -- the tensor-parallel plan of kf_config_shard_of applied to the vendor layout.  *bytes_out receives the blob size;
// out_blob may be NULL to query it.
// ---- fish.kun, the reference's own container (csrc/Tensor/KunFile.cpp) --------------------------------------------------------------
extern "C" int kf_model_save_kun(kf_model* m, const char* path) {
    if (!m || !path) return KF_ERR_BAD_ARG;
    try {
        // jsConfig of Fish::SAFETENSOR_Serialize (Serialize.cpp:912-922): {"vendor", "CLI_params": {"config": ...}, "tokenizer": {"tokens": ""}}
        const std::string cfg = "{\"vendor\":\"koifish_b200\",\"CLI_params\":{\"config\":" + m->config_text + "},\"tokenizer\":{\"tokens\":\"\"}}";
        return m->fish->SaveKun(path, cfg);
    } catch (const std::exception& e) {
        m->fish->error = std::string("save_kun: ") + e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_load_kun(kf_model* m, const char* path, int* n_loaded_out, int* n_skipped_out) {
    if (!m || !path) return KF_ERR_BAD_ARG;
    if (n_loaded_out) *n_loaded_out = 0;
    if (n_skipped_out) *n_skipped_out = 0;
    try {
        return m->fish->LoadKun(path, n_loaded_out, n_skipped_out);
    } catch (const std::exception& e) {
        m->fish->error = std::string("load_kun: ") + e.what();
        return KF_ERR_BAD_ARG;
    }
}
// host only: the header of a .kun file as JSON text [{"name","dtype","shape","szData","szGama","offset"}, ...] in file order
extern "C" int kf_kun_index(const char* path, char** json_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || !json_out) return KF_ERR_BAD_ARG;
    *json_out = nullptr;
    KunFile f;
    std::string err;
    if (kun_parse(path, &f, &err) != 0) {
        if (err_out) *err_out = dup_cstr(err);
        return KF_ERR_BAD_ARG;
    }
    std::string j = "[";
    for (size_t i = 0; i < f.entries.size(); i++) {
        const KunEntry& e = f.entries[i];
        JSON name;
        name.kind = JSON::String, name.str = e.name;
        JSON dt;
        dt.kind = JSON::String, dt.str = e.dtype;
        j += (i ? ",{\"name\":" : "{\"name\":") + json_dump(name) + ",\"dtype\":" + json_dump(dt) + ",\"shape\":[";
        for (size_t d = 0; d < e.shape.size(); d++) j += (d ? "," : "") + std::to_string(e.shape[d]);
        j += "],\"szData\":" + std::to_string(e.szData) + ",\"szGama\":" + std::to_string(e.szGama) + ",\"offset\":" + std::to_string(e.begin) + "}";
    }
    j += "]";
    *json_out = dup_cstr(j);
    return KF_OK;
}
// host only: the "__koifish__config__" entry (msgpack) as JSON text; "" when the file has none
extern "C" int kf_kun_config(const char* path, char** json_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || !json_out) return KF_ERR_BAD_ARG;
    *json_out = nullptr;
    KunFile f;
    std::string err, text;
    if (kun_parse(path, &f, &err) != 0 || kun_config_json(f, &text, &err) != 0) {
        if (err_out) *err_out = dup_cstr(err);
        return KF_ERR_BAD_ARG;
    }
    *json_out = dup_cstr(text);
    return KF_OK;
}
// host only: write a .kun from host blobs (what kf_model_save_kun does after copying every tensor off the device).  shapes: n x 2 (second 0 = vector)
extern "C" int kf_kun_write(const char* path, const char* config_json, int n, const char* const* names, const char* const* dtypes, const int64_t* shapes,
                            const uint64_t* sz_data, const uint64_t* sz_gama, const void* const* blobs, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || n < 0 || (n && (!names || !dtypes || !shapes || !sz_data || !sz_gama || !blobs))) return KF_ERR_BAD_ARG;
    std::vector<KunTensorOut> outs((size_t)n);
    for (int i = 0; i < n; i++) {
        if (!names[i] || !dtypes[i]) return KF_ERR_BAD_ARG;
        outs[i].name = names[i], outs[i].dtype = dtypes[i];
        outs[i].shape[0] = shapes[2 * i], outs[i].shape[1] = shapes[2 * i + 1];
        outs[i].szData = sz_data[i], outs[i].szGama = sz_gama[i], outs[i].blob = blobs[i];
    }
    std::string err;
    if (kun_write(path, config_json ? config_json : "", outs, &err) != 0) {
        if (err_out) *err_out = dup_cstr(err);
        return KF_ERR_BAD_ARG;
    }
    return KF_OK;
}
extern "C" int kf_config_awq_shard(const char* config_json, const char* tensor_name, int rank, int world, const void* qweight, const void* qzeros,
                                   const void* scales, void* out_blob, size_t capacity, size_t* bytes_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !tensor_name || !bytes_out) return KF_ERR_BAD_ARG;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        if (c.awq_repack) {  // gpt.awq_repack = 1: the blob is PackedQ 4-bit data || gama of the window (AwqRepackWindow)
            int sh[6];
            if (!ShardPlan(c, tensor_name, rank, world, sh, sh + 1, sh + 2, sh + 3, sh + 4, sh + 5) || sh[2] % 32 || sh[3] % 128 || sh[4] % 8 || sh[5] % 128) {
                if (err_out) *err_out = dup_cstr("unknown tensor name, or the tensor-parallel window is not in whole 32-column blocks / 128-row groups");
                return KF_ERR_BAD_ARG;
            }
            *bytes_out = AwqRepackBytes(sh[2], sh[3]);
            if (!out_blob) return KF_OK;
            if (!qweight || !qzeros || !scales || capacity < *bytes_out) return KF_ERR_BAD_ARG;
            AwqRepackWindow(qweight, qzeros, scales, sh[1], sh[0], sh[4], sh[2], sh[5], sh[3], (uint8_t*)out_blob);
            return KF_OK;
        }
        int sh[6];
        if (!ShardPlan(c, tensor_name, rank, world, sh, sh + 1, sh + 2, sh + 3, sh + 4, sh + 5) || sh[2] % 32 || sh[3] % 128 || sh[4] % 8 || sh[5] % 128) {
            if (err_out) *err_out = dup_cstr("unknown tensor name, or the tensor-parallel window is not in whole 32-column blocks / 128-row groups");
            return KF_ERR_BAD_ARG;
        }
        const int OC = sh[0], IC = sh[1], OCl = sh[2], ICl = sh[3];
        *bytes_out = (size_t)ICl * OCl / 2 + (size_t)(ICl / 128) * (OCl / 8) * 4 + (size_t)(ICl / 128) * OCl * 2;
        if (!out_blob) return KF_OK;
        if (!qweight || !qzeros || !scales || capacity < *bytes_out) return KF_ERR_BAD_ARG;
        AwqShardWindow(qweight, qzeros, scales, IC, OC, sh[4], OCl, sh[5], ICl, (uint8_t*)out_blob);
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_config_shard_of(const char* config_json, const char* tensor_name, int rank, int world, int* shape_out /* rows_g, cols_g, rows_l,
                                  cols_l, row0, col0 */, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!config_json || !tensor_name || !shape_out) return KF_ERR_BAD_ARG;
    try {
        MODEL_CARD c = MODEL_CARD::FromJSON(JSON::parse(config_json));
        if (!ShardPlan(c, tensor_name, rank, world, shape_out, shape_out + 1, shape_out + 2, shape_out + 3, shape_out + 4, shape_out + 5)) {
            if (err_out) *err_out = dup_cstr("unknown tensor name or tensor-parallel degree does not divide the model");
            return KF_ERR_BAD_ARG;
        }
        return KF_OK;
    } catch (const std::exception& e) {
        if (err_out) *err_out = dup_cstr(e.what());
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_save(kf_model* m, const char* path) {
    if (!m || !path) return KF_ERR_BAD_ARG;
    return m->fish->SaveBlobs(path);
}
extern "C" int kf_model_load(kf_model* m, const char* path) {
    if (!m || !path) return KF_ERR_BAD_ARG;
    return m->fish->LoadBlobs(path);
}
extern "C" int kf_model_set_sampler(kf_model* m, float temperature, int top_k, float top_p, uint64_t seed, int selection) {
    if (!m) return KF_ERR_BAD_ARG;
    return m->fish->SetSampler(temperature, top_k, top_p, seed, selection);
}
extern "C" int kf_model_generate(kf_model* m, const int32_t* prompt_ids, int n_prompt, int pos0, int max_new_tokens, int eos_id, int32_t* out_ids, int* n_out,
                                 int* stop_reason_out) {
    if (!m) return KF_ERR_BAD_ARG;
    try {
        return m->fish->Generate(prompt_ids, n_prompt, pos0, max_new_tokens, eos_id, out_ids, n_out, stop_reason_out);
    } catch (const std::exception& e) {
        m->fish->error = e.what();
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_model_set_graphs(kf_model* m, int enable) {
    if (!m) return KF_ERR_BAD_ARG;
    m->fish->use_graphs = enable != 0;
    if (!enable) m->fish->ResetGraphs();
    return KF_OK;
}

// ---- HF safetensors (Fish::LoadFolderOfST, reference src/Manifold/Serialize.cpp:1010-1100) ------------------------------------------
// index of one file as JSON text: [{"name","dtype","shape","nbytes"}, ...] in file order.  Host only: no device, no model.
extern "C" int kf_safetensors_index(const char* path, char** json_out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || !json_out) return KF_ERR_BAD_ARG;
    *json_out = nullptr;
    KfStFile f;
    std::string err;
    if (kf_st_parse(path, &f, &err) != 0) {
        if (err_out) *err_out = dup_cstr(err);
        return KF_ERR_BAD_ARG;
    }
    std::string j = "[";
    for (size_t i = 0; i < f.entries.size(); i++) {
        const KfStEntry& e = f.entries[i];
        j += (i ? ",{\"name\":\"" : "{\"name\":\"") + e.name + "\",\"dtype\":\"" + e.dtype + "\",\"shape\":[";
        for (size_t d = 0; d < e.shape.size(); d++) j += (d ? "," : "") + std::to_string(e.shape[d]);
        j += "],\"nbytes\":" + std::to_string(e.end - e.begin) + "}";
    }
    j += "]";
    *json_out = dup_cstr(j);
    return KF_OK;
}
// one tensor of one file as bf16 (BF16 / F16 / F32 sources, round to nearest even), host only: out_bf16 holds `capacity` elements
extern "C" int kf_safetensors_read_bf16(const char* path, const char* name, void* out_bf16, size_t capacity, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || !name || !out_bf16) return KF_ERR_BAD_ARG;
    KfStFile f;
    std::string err;
    int rc = kf_st_parse(path, &f, &err);
    if (!rc) {
        rc = -1, err = std::string("'") + path + "': no tensor '" + name + "'";
        for (const KfStEntry& e : f.entries) {
            if (e.name != name) continue;
            size_t n = 1;
            for (int64_t d : e.shape) n *= (size_t)d;
            if (e.dtype != "BF16" && e.dtype != "F16" && e.dtype != "F32") {
                err = "'" + e.name + "': dtype " + e.dtype + " cannot be read as bf16";
            } else if (n > capacity) {
                err = "'" + e.name + "': " + std::to_string(n) + " elements, buffer holds " + std::to_string(capacity);
            } else {
                std::vector<uint8_t> raw(e.end - e.begin);
                rc = kf_st_read(f, e, raw.data(), &err);
                if (!rc) kf_st_to_bf16(e.dtype, raw.data(), n, (uint16_t*)out_bf16);
            }
            break;
        }
    }
    if (rc && err_out) *err_out = dup_cstr(err);
    return rc ? KF_ERR_BAD_ARG : KF_OK;
}
// Every tensor of the file (or of every *.safetensors of the directory) whose name the model knows is set from it -- BF16 / F16 / F32
// sources, rounded to bf16, sharded for this rank and quantised per the quantizer card exactly as kf_model_set_tensor does.  Names the
// model does not have (rotary inv_freq, a tied lm_head.weight, biases ...) are skipped and counted.
// Vendor-quantised linears (an AWQ checkpoint, e.g. Qwen3-32B-AWQ: <prefix>.qweight I32 [in][out / 8], <prefix>.qzeros I32 [in / 128][out / 8],
// <prefix>.scales F16 [in / 128][out]; reference GeQuant::ExTensor GeQuant.cpp:144-200, src/Python/test_awq.py:33-71) become the model's
// <prefix>.weight in the AWQ layout (kf_model_set_tensor_awq) once all three arrays have been seen -- they may sit in different shards;
// a triple left incomplete at the end is an error.  Each complete triple counts as one loaded tensor.
namespace {
struct AwqPart {
    std::vector<uint8_t> bytes;
    std::vector<int64_t> shape;
    bool seen = false;
};
struct AwqTriple {
    AwqPart qweight, qzeros, scales;
};
}  // namespace
extern "C" int kf_model_load_safetensors(kf_model* m, const char* path_or_dir, int* n_loaded_out, int* n_skipped_out) {
    if (!m || !path_or_dir) return KF_ERR_BAD_ARG;
    int loaded = 0, skipped = 0;
    if (n_loaded_out) *n_loaded_out = 0;
    if (n_skipped_out) *n_skipped_out = 0;
    try {
        std::vector<std::string> files;
        std::string err;
        if (kf_st_list(path_or_dir, &files, &err) != 0) throw std::runtime_error(err);
        std::vector<uint8_t> raw;
        std::vector<uint16_t> bf;
        std::map<std::string, AwqTriple> awq;  // by <prefix>
        for (const std::string& path : files) {
            KfStFile f;
            if (kf_st_parse(path, &f, &err) != 0) throw std::runtime_error(err);
            for (const KfStEntry& e : f.entries) {
                auto ends_with = [&](const char* suf) {
                    const size_t n = strlen(suf);
                    return e.name.size() > n && e.name.compare(e.name.size() - n, n, suf) == 0;
                };
                const int part = ends_with(".qweight") ? 0 : ends_with(".qzeros") ? 1 : ends_with(".scales") ? 2 : -1;
                if (part >= 0) {
                    const std::string prefix = e.name.substr(0, e.name.rfind('.'));
                    const std::string wname  = prefix + ".weight";
                    if (!m->fish->GetTensor(wname)) {
                        skipped++;
                        continue;
                    }
                    if (e.dtype != (part == 2 ? "F16" : "I32") || e.shape.size() != 2)
                        throw std::runtime_error("'" + e.name + "': the AWQ layout stores qweight / qzeros as 2-D I32 and scales as 2-D F16, found " + e.dtype);
                    AwqTriple& t = awq[prefix];
                    AwqPart& p   = part == 0 ? t.qweight : part == 1 ? t.qzeros : t.scales;
                    if (p.seen) throw std::runtime_error("'" + e.name + "' appears twice");
                    p.bytes.resize(e.end - e.begin);
                    if (kf_st_read(f, e, p.bytes.data(), &err) != 0) throw std::runtime_error(err);
                    p.shape = e.shape, p.seen = true;
                    if (t.qweight.seen && t.qzeros.seen && t.scales.seen) {
                        const int64_t IC = t.qweight.shape[0], OC = t.scales.shape[1];
                        if (IC <= 0 || OC <= 0 || IC > 0x7fffffff || OC > 0x7fffffff || IC % 128 || OC % 8 || t.qweight.shape[1] != OC / 8 ||
                            t.qzeros.shape[0] != IC / 128 || t.qzeros.shape[1] != OC / 8 || t.scales.shape[0] != IC / 128)
                            throw std::runtime_error("'" + prefix + "': qweight / qzeros / scales shapes are not [in][out/8], [in/128][out/8], [in/128][out] "
                                                     "(4-bit codes, group_size 128)");
                        const int rc = m->fish->SetTensorAWQ(wname, t.qweight.bytes.data(), t.qzeros.bytes.data(), t.scales.bytes.data(), (int)IC, (int)OC);
                        if (rc) return rc;  // Fish::error is set
                        awq.erase(prefix);
                        loaded++;
                    }
                    continue;
                }
                if (!m->fish->GetTensor(e.name)) {
                    skipped++;
                    continue;
                }
                if (e.dtype != "BF16" && e.dtype != "F16" && e.dtype != "F32") throw std::runtime_error("'" + e.name + "': dtype " + e.dtype + " cannot be a weight");
                if (e.shape.empty() || e.shape.size() > 2) throw std::runtime_error("'" + e.name + "': expected a vector or a matrix");
                const int64_t rows = e.shape.size() == 2 ? e.shape[0] : 1, cols = e.shape.back();
                if (rows <= 0 || cols <= 0 || rows > 0x7fffffff || cols > 0x7fffffff) throw std::runtime_error("'" + e.name + "': bad shape");
                raw.resize(e.end - e.begin);
                if (kf_st_read(f, e, raw.data(), &err) != 0) throw std::runtime_error(err);
                bf.resize((size_t)rows * cols);
                kf_st_to_bf16(e.dtype, raw.data(), bf.size(), bf.data());
                const int rc = m->fish->SetTensor(e.name, bf.data(), (int)rows, (int)cols);
                if (rc) return rc;  // Fish::error is set
                loaded++;
            }
        }
        if (!awq.empty())
            throw std::runtime_error("'" + awq.begin()->first + "': incomplete AWQ tensor (needs .qweight, .qzeros and .scales)");
    } catch (const std::exception& e) {
        m->fish->error = std::string("load_safetensors: ") + e.what();
        return KF_ERR_BAD_ARG;
    }
    if (n_loaded_out) *n_loaded_out = loaded;
    if (n_skipped_out) *n_skipped_out = skipped;
    return KF_OK;
}
