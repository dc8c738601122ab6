// kf_tokenizer_api.cpp -- extern "C" surface of csrc/TokenSet (include/kf_tokenizer.h).  Exceptions never cross the boundary.
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "HF_Tokenizer.hpp"
#include "kf_device.h"
#include "kf_tokenizer.h"

using namespace koifish;

struct kf_tokenizer {
    std::shared_ptr<HF_Tokenizer> tk;
};
static char* dup_str(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
static std::string json_quote(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') {
            o += '\\', o += (char)c;
        } else if (c < 0x20) {
            char b[8];
            snprintf(b, sizeof(b), "\\u%04x", c);
            o += b;
        } else
            o += (char)c;
    }
    return o + "\"";
}
static int wrap(std::shared_ptr<HF_Tokenizer> tk, const std::string& err, kf_tokenizer** out, char** err_out) {
    if (!tk) {
        if (err_out) *err_out = dup_str(err);
        return KF_ERR_UNSUPPORTED;
    }
    *out = new kf_tokenizer{tk};
    return KF_OK;
}
extern "C" int kf_tokenizer_load(const char* path, kf_tokenizer** out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!path || !out) return KF_ERR_BAD_ARG;
    *out = nullptr;
    std::string err;
    return wrap(HF_Tokenizer::FromPath(path, &err), err, out, err_out);
}
extern "C" int kf_tokenizer_from_json(const char* text, const char* cfg, kf_tokenizer** out, char** err_out) {
    if (err_out) *err_out = nullptr;
    if (!text || !out) return KF_ERR_BAD_ARG;
    *out = nullptr;
    std::string err;
    return wrap(HF_Tokenizer::FromJSONText(text, cfg ? cfg : "", &err), err, out, err_out);
}
extern "C" int kf_tokenizer_destroy(kf_tokenizer* t) {
    delete t;
    return KF_OK;
}
extern "C" int kf_tokenizer_encode(const kf_tokenizer* t, const char* text, size_t nbytes, int32_t* ids, size_t capacity, size_t* n_out) {
    if (!t || (!text && nbytes) || !n_out) return KF_ERR_BAD_ARG;
    try {
        const std::vector<int> v = t->tk->encode(std::string(text ? text : "", nbytes));
        *n_out = v.size();
        if (!ids) return KF_OK;
        if (v.size() > capacity) return KF_ERR_BAD_ARG;
        for (size_t i = 0; i < v.size(); i++) ids[i] = v[i];
        return KF_OK;
    } catch (const std::exception&) {
        *n_out = 0;
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_tokenizer_decode(const kf_tokenizer* t, const int32_t* ids, size_t n, int skip_special, char** text_out) {
    if (!t || (!ids && n) || !text_out) return KF_ERR_BAD_ARG;
    try {
        *text_out = dup_str(t->tk->decode(std::vector<int>(ids, ids + n), skip_special != 0));
        return *text_out ? KF_OK : KF_ERR_OOM;
    } catch (const std::exception&) {
        return KF_ERR_BAD_ARG;
    }
}
struct kf_decode_stream {
    std::string pending;
};
extern "C" int kf_decode_stream_create(kf_decode_stream** out) {
    if (!out) return KF_ERR_BAD_ARG;
    *out = new kf_decode_stream();
    return KF_OK;
}
extern "C" int kf_decode_stream_destroy(kf_decode_stream* s) {
    delete s;
    return KF_OK;
}
extern "C" int kf_decode_stream_push(const kf_tokenizer* t, kf_decode_stream* s, int id, int skip_special, char** text_out) {
    if (!t || !s || !text_out) return KF_ERR_BAD_ARG;
    *text_out = dup_str(t->tk->stream_push(&s->pending, id, skip_special != 0));
    return *text_out ? KF_OK : KF_ERR_OOM;
}
extern "C" int kf_decode_stream_flush(kf_decode_stream* s, char** text_out) {
    if (!s || !text_out) return KF_ERR_BAD_ARG;
    *text_out = dup_str(HF_Tokenizer::stream_flush(&s->pending));
    return *text_out ? KF_OK : KF_ERR_OOM;
}
extern "C" int kf_tokenizer_token_to_id(const kf_tokenizer* t, const char* token) { return t && token ? t->tk->token_to_id(token) : -1; }
extern "C" int kf_tokenizer_id_to_token(const kf_tokenizer* t, int id, char** out) {
    if (!t || !out) return KF_ERR_BAD_ARG;
    if (id < 0 || id >= t->tk->vocab_size()) return KF_ERR_BAD_ARG;
    *out = dup_str(t->tk->id_to_token(id));
    return *out ? KF_OK : KF_ERR_OOM;
}
extern "C" int kf_tokenizer_vocab_size(const kf_tokenizer* t) { return t ? t->tk->vocab_size() : 0; }
extern "C" int kf_tokenizer_eos_id(const kf_tokenizer* t) { return t ? t->tk->eos_token_id() : -1; }
extern "C" int kf_tokenizer_bos_id(const kf_tokenizer* t) { return t ? t->tk->bos_token_id() : -1; }
extern "C" int kf_tokenizer_pad_id(const kf_tokenizer* t) { return t ? t->tk->pad_token_id() : -1; }
extern "C" int kf_tokenizer_is_special(const kf_tokenizer* t, int id) { return t && t->tk->is_special(id) ? 1 : 0; }
extern "C" int kf_text_nfc(const char* text, size_t nbytes, char** out) {
    if ((!text && nbytes) || !out) return KF_ERR_BAD_ARG;
    try {
        *out = dup_str(HF_Tokenizer::NFC(std::string(text ? text : "", nbytes)));
        return *out ? KF_OK : KF_ERR_OOM;
    } catch (const std::exception&) {
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_tokenizer_pre_tokenize(const kf_tokenizer* t, const char* text, size_t nbytes, char** out) {
    if (!t || (!text && nbytes) || !out) return KF_ERR_BAD_ARG;
    try {
        std::string j = "[";
        bool first    = true;
        for (const std::string& p : t->tk->pre_tokenize(std::string(text ? text : "", nbytes))) {
            j += (first ? "" : ",") + json_quote(p);
            first = false;
        }
        *out = dup_str(j + "]");
        return *out ? KF_OK : KF_ERR_OOM;
    } catch (const std::exception&) {
        return KF_ERR_BAD_ARG;
    }
}
extern "C" int kf_chatml_prompt(const char* system, const char* user, int enable_thinking, char** out) {
    if (!user || !out) return KF_ERR_BAD_ARG;
    *out = dup_str(ChatMLPrompt(system ? system : "", user, enable_thinking != 0));
    return *out ? KF_OK : KF_ERR_OOM;
}
extern "C" int kf_chatml_render(const char* const* roles, const char* const* contents, int n, int enable_thinking, char** out) {
    if (n < 0 || (n && (!roles || !contents)) || !out) return KF_ERR_BAD_ARG;
    std::vector<std::pair<std::string, std::string>> lines;
    for (int i = 0; i < n; i++) {
        if (!roles[i] || !contents[i]) return KF_ERR_BAD_ARG;
        lines.emplace_back(roles[i], contents[i]);
    }
    *out = dup_str(ChatMLRender(lines, enable_thinking != 0));
    return *out ? KF_OK : KF_ERR_OOM;
}
