// ref_tokenizer.cpp -- TEST INFRASTRUCTURE ONLY.  The reference's own tokenizer (src/TokenSet/HF_Tokenizer.cpp + Dictionary.cpp, with the oniguruma and
// utf8proc sources it vendors under src/Utils) compiled where it lies into oracle/_ref/libkoifish_reftok.so (`make -C oracle reftok`; oniguruma's
// cmake-generated config.h is replaced by oracle/onig_config/config.h), so that csrc/TokenSet can be compared with it on the same tokenizer.json.
// The rest of the framework those files mention is bound to 0 at link time and never reached.  Nothing of the reference is copied.
#include <cstring>
#include <string>
#include <vector>
#include "TokenSet/Dictionary.hpp"
extern "C" void* reftok_load(const char* json_text) {
    auto* tk = new HF_Tokenizer();
    if (!tk->load_from_json_str(json_text)) { return nullptr; }
    return tk;
}
extern "C" int reftok_encode(void* h, const char* text, int* ids, int cap) {
    auto* tk = (HF_Tokenizer*)h;
    std::vector<int> v = tk->encode(text, false);
    if ((int)v.size() > cap) return -(int)v.size();
    for (size_t i = 0; i < v.size(); i++) ids[i] = v[i];
    return (int)v.size();
}
extern "C" int reftok_decode(void* h, const int* ids, int n, int skip_special, char* out, int cap) {
    auto* tk = (HF_Tokenizer*)h;
    std::string s = tk->decode(std::vector<int>(ids, ids + n), skip_special != 0);
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
