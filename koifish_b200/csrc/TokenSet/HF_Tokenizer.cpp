// HF_Tokenizer.cpp -- see HF_Tokenizer.hpp.  Algorithms restated from the published behaviour of the HF `tokenizers` crate (v0.22, the version
// the golden vectors under tests/golden/tokenizer/ were produced with) which the reference's src/TokenSet/HF_Tokenizer.cpp ports:
//   AddedVocabulary::extract_and_normalize  -> split_added            (reference HF_Tokenizer.cpp:1420-1520)
//   NFC normalizer (unicode-normalization)  -> HF_Tokenizer::NFC       (reference class NFKCNormalizer :202 is its only normalisation form)
//   Split{Regex, Isolated} + ByteLevel      -> pre_tokenize + kByteChar (reference :375-420, :480-528)
//   BPE::merge_word / Word::merge_all       -> bpe_word                (reference BPEModel :586-716)
//   ByteLevel decoder + from_utf8_lossy     -> decode                  (reference ByteLevelDecoder :1028-1060)
#include "HF_Tokenizer.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <queue>
#include <stdexcept>

#include "../Utils/json_lite.hpp"
#include "unicode_tables.hpp"

namespace koifish {
namespace {

// ------------------------------------------------------------------------------------------------ UTF-8
bool utf8_decode(const std::string& s, std::vector<uint32_t>* out) {
    out->clear();
    out->reserve(s.size());
    const unsigned char* p = (const unsigned char*)s.data();
    const size_t n         = s.size();
    for (size_t i = 0; i < n;) {
        const unsigned c = p[i];
        uint32_t cp;
        int len;
        if (c < 0x80)
            cp = c, len = 1;
        else if (c >= 0xC2 && c <= 0xDF)
            cp = c & 0x1F, len = 2;
        else if (c >= 0xE0 && c <= 0xEF)
            cp = c & 0x0F, len = 3;
        else if (c >= 0xF0 && c <= 0xF4)
            cp = c & 0x07, len = 4;
        else
            return false;
        if (i + len > n) return false;
        for (int k = 1; k < len; k++) {
            if ((p[i + k] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (p[i + k] & 0x3F);
        }
        if ((len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) || (len == 4 && (cp < 0x10000 || cp > 0x10FFFF))) return false;
        out->push_back(cp);
        i += len;
    }
    return true;
}
void utf8_append(std::string* s, uint32_t cp) {
    if (cp < 0x80)
        *s += (char)cp;
    else if (cp < 0x800)
        *s += (char)(0xC0 | (cp >> 6)), *s += (char)(0x80 | (cp & 0x3F));
    else if (cp < 0x10000)
        *s += (char)(0xE0 | (cp >> 12)), *s += (char)(0x80 | ((cp >> 6) & 0x3F)), *s += (char)(0x80 | (cp & 0x3F));
    else
        *s += (char)(0xF0 | (cp >> 18)), *s += (char)(0x80 | ((cp >> 12) & 0x3F)), *s += (char)(0x80 | ((cp >> 6) & 0x3F)), *s += (char)(0x80 | (cp & 0x3F));
}
// String::from_utf8_lossy: every maximal ill-formed prefix becomes one U+FFFD.  hold_tail: a sequence that is well-formed so far but cut by the
// end of the buffer is left alone (*consumed stops in front of it) -- the streaming decoder waits for the bytes of the next token.
std::string utf8_lossy_prefix(const std::string& s, bool hold_tail, size_t* consumed) {
    std::string out;
    const unsigned char* p = (const unsigned char*)s.data();
    const size_t n         = s.size();
    size_t i               = 0;
    while (i < n) {
        const unsigned c = p[i];
        if (c < 0x80) {
            out += (char)c, i++;
            continue;
        }
        int len = 0;
        unsigned lo = 0x80, hi = 0xBF;  // allowed range of the SECOND byte (Unicode table 3-7)
        if (c >= 0xC2 && c <= 0xDF)
            len = 2;
        else if (c >= 0xE0 && c <= 0xEF)
            len = 3, lo = c == 0xE0 ? 0xA0 : 0x80, hi = c == 0xED ? 0x9F : 0xBF;
        else if (c >= 0xF0 && c <= 0xF4)
            len = 4, lo = c == 0xF0 ? 0x90 : 0x80, hi = c == 0xF4 ? 0x8F : 0xBF;
        if (!len) {
            out += "\xEF\xBF\xBD", i++;
            continue;
        }
        size_t k = 1;
        bool cut = false;
        for (; k < (size_t)len; k++) {
            if (i + k >= n) {
                cut = true;
                break;
            }
            const unsigned b = p[i + k];
            if (k == 1 ? (b < lo || b > hi) : ((b & 0xC0) != 0x80)) break;
        }
        if (cut && hold_tail) break;
        if (k == (size_t)len)
            out.append(s, i, len);
        else
            out += "\xEF\xBF\xBD";
        i += k;
    }
    if (consumed) *consumed = i;
    return out;
}
std::string utf8_lossy(const std::string& s) { return utf8_lossy_prefix(s, false, nullptr); }

// ------------------------------------------------------------------------------------------------ Unicode properties
bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (cp < r[mid][0])
            hi = mid - 1;
        else if (cp > r[mid][1])
            lo = mid + 1;
        else
            return true;
    }
    return false;
}
inline bool isL(uint32_t c) {
    if (c < 0x80) return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
    return in_ranges(kf_unicode::kLetter, kf_unicode::kLetter_N, c);
}
inline bool isN(uint32_t c) {
    if (c < 0x80) return c >= '0' && c <= '9';
    return in_ranges(kf_unicode::kNumber, kf_unicode::kNumber_N, c);
}
inline bool isS(uint32_t c) {
    if (c < 0x80) return (c >= 9 && c <= 13) || c == 32;
    return in_ranges(kf_unicode::kWhiteSpace, kf_unicode::kWhiteSpace_N, c);
}
int ccc_of(uint32_t cp) {
    if (cp < 0x300) return 0;
    int lo = 0, hi = kf_unicode::kCCC_N - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (cp < kf_unicode::kCCC[mid][0])
            hi = mid - 1;
        else if (cp > kf_unicode::kCCC[mid][0])
            lo = mid + 1;
        else
            return (int)kf_unicode::kCCC[mid][1];
    }
    return 0;
}
const uint32_t* decomp_of(uint32_t cp) {
    int lo = 0, hi = kf_unicode::kDecomp_N - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (cp < kf_unicode::kDecomp[mid][0])
            hi = mid - 1;
        else if (cp > kf_unicode::kDecomp[mid][0])
            lo = mid + 1;
        else
            return kf_unicode::kDecomp[mid];
    }
    return nullptr;
}
constexpr uint32_t SBase = 0xAC00, LBase = 0x1100, VBase = 0x1161, TBase = 0x11A7, LCount = 19, VCount = 21, TCount = 28, NCount = VCount * TCount,
                   SCount = LCount * NCount;
uint32_t compose_pair(uint32_t a, uint32_t b) {
    if (a >= LBase && a < LBase + LCount && b >= VBase && b < VBase + VCount) return SBase + ((a - LBase) * VCount + (b - VBase)) * TCount;
    if (a >= SBase && a < SBase + SCount && (a - SBase) % TCount == 0 && b > TBase && b < TBase + TCount) return a + (b - TBase);
    int lo = 0, hi = kf_unicode::kComp_N - 1;
    while (lo <= hi) {
        const int mid       = (lo + hi) / 2;
        const uint32_t* e = kf_unicode::kComp[mid];
        if (a < e[0] || (a == e[0] && b < e[1]))
            hi = mid - 1;
        else if (a > e[0] || (a == e[0] && b > e[1]))
            lo = mid + 1;
        else
            return e[2];
    }
    return 0;
}
void decompose_into(uint32_t cp, std::vector<uint32_t>* out) {
    if (cp >= SBase && cp < SBase + SCount) {  // Hangul syllables decompose arithmetically
        const uint32_t s = cp - SBase;
        out->push_back(LBase + s / NCount);
        out->push_back(VBase + (s % NCount) / TCount);
        if (s % TCount) out->push_back(TBase + s % TCount);
        return;
    }
    const uint32_t* d = cp < 0xC0 ? nullptr : decomp_of(cp);
    if (!d) {
        out->push_back(cp);
        return;
    }
    decompose_into(d[1], out);
    if (d[2]) decompose_into(d[2], out);
}

// GPT-2 bytes_to_unicode: printable Latin-1 bytes map to themselves, the other 68 to U+0100 ...
struct ByteChars {
    uint32_t cp[256];
    std::string utf8[256];
    std::unordered_map<uint32_t, int> back;
    ByteChars() {
        int n = 0;
        for (int b = 0; b < 256; b++) {
            const bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            cp[b]           = keep ? (uint32_t)b : (uint32_t)(256 + n++);
            utf8_append(&utf8[b], cp[b]);
            back[cp[b]] = b;
        }
    }
};
const ByteChars& byte_chars() {
    static const ByteChars t;
    return t;
}

std::string read_file(const std::string& path, bool* ok) {
    FILE* f = fopen(path.c_str(), "rb");
    *ok     = f != nullptr;
    std::string s;
    if (!f) return s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}
// the pre-tokenisation patterns this scanner implements (tokenizer.json "pre_tokenizer" -> Split -> pattern.Regex)
const char* kPatternQwen   = "(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\\r\\n\\p{L}\\p{N}]?\\p{L}+|\\p{N}| ?[^\\s\\p{L}\\p{N}]+[\\r\\n]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+";
const char* kPatternLlama3 = "(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\\r\\n\\p{L}\\p{N}]?\\p{L}+|\\p{N}{1,3}| ?[^\\s\\p{L}\\p{N}]+[\\r\\n]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+";

uint32_t fold(uint32_t c) {  // the case folding (?i:...) applies to the seven contraction letters
    if (c >= 'A' && c <= 'Z') return c + 32;
    if (c == 0x17F) return 's';   // LATIN SMALL LETTER LONG S folds to s
    if (c == 0x212A) return 'k';  // KELVIN SIGN (not a contraction letter; kept for completeness)
    return c;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ NFC
std::string HF_Tokenizer::NFC(const std::string& utf8) {
    bool plain = true;  // nothing below U+0300 decomposes into, or composes with, anything under NFC
    for (unsigned char c : utf8)
        if (c >= 0xCC) {  // first byte of U+0300 and above
            plain = false;
            break;
        }
    if (plain) return utf8;
    std::vector<uint32_t> in, d;
    if (!utf8_decode(utf8, &in)) throw std::runtime_error("tokenizer: text is not valid UTF-8");
    d.reserve(in.size() + 8);
    for (uint32_t cp : in) decompose_into(cp, &d);
    // canonical ordering: runs of non-starters sorted by combining class, stably
    for (size_t i = 0; i < d.size();) {
        if (ccc_of(d[i]) == 0) {
            i++;
            continue;
        }
        size_t j = i;
        while (j < d.size() && ccc_of(d[j]) != 0) j++;
        std::stable_sort(d.begin() + i, d.begin() + j, [](uint32_t a, uint32_t b) { return ccc_of(a) < ccc_of(b); });
        i = j;
    }
    // canonical composition
    std::vector<uint32_t> o;
    o.reserve(d.size());
    long starter = -1;
    int last_ccc = 0;
    for (uint32_t c : d) {
        const int cc = ccc_of(c);
        if (starter >= 0) {
            const bool adjacent = (long)o.size() - 1 == starter;
            if (adjacent || (last_ccc != 0 && last_ccc < cc)) {  // not blocked from the starter
                if (const uint32_t comp = compose_pair(o[starter], c)) {
                    o[starter] = comp;
                    continue;
                }
            }
        }
        if (cc == 0) starter = (long)o.size();
        last_ccc = cc;
        o.push_back(c);
    }
    std::string out;
    out.reserve(utf8.size());
    for (uint32_t cp : o) utf8_append(&out, cp);
    return out;
}

// ------------------------------------------------------------------------------------------------ pre-tokenisation
// One pass of the pattern over the code points; alternatives tried in order at every position, as a backtracking engine does.
std::vector<std::string> HF_Tokenizer::pre_tokenize(const std::string& utf8) const {
    std::vector<uint32_t> cp;
    if (!utf8_decode(utf8, &cp)) throw std::runtime_error("tokenizer: text is not valid UTF-8");
    std::vector<size_t> off(cp.size() + 1, 0);  // byte offset of every code point
    for (size_t i = 0; i < cp.size(); i++) off[i + 1] = off[i] + (cp[i] < 0x80 ? 1 : cp[i] < 0x800 ? 2 : cp[i] < 0x10000 ? 3 : 4);
    std::vector<std::string> pieces;
    const size_t n = cp.size();
    size_t i = 0;
    auto emit = [&](size_t j) {
        pieces.emplace_back(utf8, off[i], off[j] - off[i]);
        i = j;
    };
    while (i < n) {
        const uint32_t c = cp[i];
        // (?i:'s|'t|'re|'ve|'m|'ll|'d)
        if (c == '\'' && i + 1 < n) {
            const uint32_t a = fold(cp[i + 1]);
            if (a == 's' || a == 't' || a == 'm' || a == 'd') {
                emit(i + 2);
                continue;
            }
            if (i + 2 < n) {
                const uint32_t b = fold(cp[i + 2]);
                if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) {
                    emit(i + 3);
                    continue;
                }
            }
        }
        // [^\r\n\p{L}\p{N}]?\p{L}+
        {
            size_t j = i;
            if (!isL(c) && c != '\r' && c != '\n' && !isN(c) && i + 1 < n && isL(cp[i + 1])) j = i + 1;
            if (isL(cp[j])) {
                while (j < n && isL(cp[j])) j++;
                emit(j);
                continue;
            }
        }
        // \p{N}  or  \p{N}{1,3}
        if (isN(c)) {
            size_t j = i;
            while (j < n && (int)(j - i) < max_digits_ && isN(cp[j])) j++;
            emit(j);
            continue;
        }
        //  ?[^\s\p{L}\p{N}]+[\r\n]*
        {
            size_t j = i + (c == ' ' ? 1 : 0), k = j;
            while (k < n && !isS(cp[k]) && !isL(cp[k]) && !isN(cp[k])) k++;
            if (k > j) {
                while (k < n && (cp[k] == '\r' || cp[k] == '\n')) k++;
                emit(k);
                continue;
            }
        }
        // whitespace: \s*[\r\n]+  |  \s+(?!\S)  |  \s+
        size_t e = i;
        while (e < n && isS(cp[e])) e++;
        if (e == i) {  // unreachable: every code point is a letter, a number, white space or "other"
            emit(i + 1);
            continue;
        }
        size_t nl = (size_t)-1;
        for (size_t k = i; k < e; k++)
            if (cp[k] == '\r' || cp[k] == '\n') nl = k;
        if (nl != (size_t)-1)
            emit(nl + 1);  // the greedy \s* gives back until [\r\n]+ can match: up to the last line break of the run
        else if (e == n || e - i == 1)
            emit(e);       // run to the end of the text, or a single white space before a word (\s+)
        else
            emit(e - 1);   // \s+(?!\S): leave the last white space to the word that follows
    }
    return pieces;
}

// ------------------------------------------------------------------------------------------------ BPE
void HF_Tokenizer::bpe_word(const std::string& piece, std::vector<int>* out) const {
    const ByteChars& bc = byte_chars();
    const size_t n      = piece.size();
    if (n == 0) return;
    if (ignore_merges_) {  // the whole pre-token is a vocabulary entry
        std::string mapped;
        for (unsigned char b : piece) mapped += bc.utf8[b];
        auto it = tok2id_.find(mapped);
        if (it != tok2id_.end()) {
            out->push_back(it->second);
            return;
        }
    }
    struct Sym {
        int id, prev, next;
        bool alive;
    };
    std::vector<Sym> sym(n);
    for (size_t i = 0; i < n; i++) sym[i] = {byte_tok_[(unsigned char)piece[i]], (int)i - 1, i + 1 < n ? (int)i + 1 : -1, true};
    struct Cand {
        int rank, pos, new_id;
        bool operator<(const Cand& o) const { return rank != o.rank ? rank > o.rank : pos > o.pos; }  // min-heap on (rank, pos)
    };
    std::priority_queue<Cand> q;
    auto push = [&](int l) {
        const int r = sym[l].next;
        if (r < 0) return;
        auto it = merges_.find(((uint64_t)(uint32_t)sym[l].id << 32) | (uint32_t)sym[r].id);
        if (it != merges_.end()) q.push({it->second.first, l, it->second.second});
    };
    for (size_t i = 0; i + 1 < n; i++) push((int)i);
    while (!q.empty()) {
        const Cand c = q.top();
        q.pop();
        if (!sym[c.pos].alive) continue;
        const int r = sym[c.pos].next;
        if (r < 0) continue;
        auto it = merges_.find(((uint64_t)(uint32_t)sym[c.pos].id << 32) | (uint32_t)sym[r].id);
        if (it == merges_.end() || it->second.second != c.new_id || it->second.first != c.rank) continue;  // the pair changed since it was queued
        sym[c.pos].id   = c.new_id;
        sym[r].alive    = false;
        sym[c.pos].next = sym[r].next;
        if (sym[r].next >= 0) sym[sym[r].next].prev = c.pos;
        if (sym[c.pos].prev >= 0) push(sym[c.pos].prev);
        push(c.pos);
    }
    for (int i = 0; i >= 0; i = sym[i].next) out->push_back(sym[i].id);
}
void HF_Tokenizer::encode_plain(const std::string& text, std::vector<int>* out) const {
    for (const std::string& piece : pre_tokenize(text)) bpe_word(piece, out);
}

// AddedVocabulary: literal matches, leftmost first, the longest token at a position; everything between goes on with id -1
void HF_Tokenizer::split_added(const std::string& text, bool normalized_pass, std::vector<std::pair<std::string, int>>* parts) const {
    size_t start = 0, i = 0;
    const size_t n = text.size();
    while (i < n) {
        const AddedToken* best = nullptr;
        for (const AddedToken& t : added_) {
            if (t.normalized != normalized_pass || t.content.empty() || t.content[0] != text[i]) continue;
            if (t.content.size() <= n - i && text.compare(i, t.content.size(), t.content) == 0 && (!best || t.content.size() > best->content.size())) best = &t;
        }
        if (!best) {
            i++;
            continue;
        }
        if (i > start) parts->emplace_back(text.substr(start, i - start), -1);
        parts->emplace_back(best->content, best->id);
        i += best->content.size();
        start = i;
    }
    if (start < n) parts->emplace_back(text.substr(start), -1);
}
std::vector<int> HF_Tokenizer::encode(const std::string& text) const {
    std::vector<int> ids;
    std::vector<std::pair<std::string, int>> raw, norm;
    split_added(text, false, &raw);  // tokens with "normalized": false are found in the text as given
    for (auto& part : raw) {
        if (part.second >= 0) {
            ids.push_back(part.second);
            continue;
        }
        const std::string t = nfc_ ? NFC(part.first) : part.first;
        if (!nfc_) {
            std::vector<uint32_t> chk;
            if (!utf8_decode(t, &chk)) throw std::runtime_error("tokenizer: text is not valid UTF-8");
        }
        norm.clear();
        split_added(t, true, &norm);  // the others after normalisation
        for (auto& p2 : norm) {
            if (p2.second >= 0)
                ids.push_back(p2.second);
            else
                encode_plain(p2.first, &ids);
        }
    }
    return ids;
}

// ------------------------------------------------------------------------------------------------ decode
std::string HF_Tokenizer::token_bytes(int id, bool skip_special) const {
    if (id < 0 || id >= (int)id2tok_.size()) return std::string();
    if (skip_special && special_[id]) return std::string();
    const ByteChars& bc    = byte_chars();
    const std::string& tok = id2tok_[id];
    // a token whose characters are all byte-level characters is those bytes; anything else stands for itself
    std::vector<uint32_t> cps;
    bool ok = utf8_decode(tok, &cps);
    std::string b;
    for (size_t i = 0; ok && i < cps.size(); i++) {
        auto it = bc.back.find(cps[i]);
        if (it == bc.back.end())
            ok = false;
        else
            b += (char)it->second;
    }
    return ok ? b : tok;
}
std::string HF_Tokenizer::decode(const std::vector<int>& ids, bool skip_special) const {
    std::string bytes;
    for (int id : ids) bytes += token_bytes(id, skip_special);
    return utf8_lossy(bytes);
}
// Streaming: the text that is certain after one more token.  A multi-byte character split over several tokens (byte-level BPE does that to rare
// characters and emoji) comes out whole with the token that completes it instead of as U+FFFD pieces; the concatenation of every push and the
// final flush equals decode() of the whole sequence.
std::string HF_Tokenizer::stream_push(std::string* pending, int id, bool skip_special) const {
    *pending += token_bytes(id, skip_special);
    size_t used = 0;
    std::string out = utf8_lossy_prefix(*pending, true, &used);
    pending->erase(0, used);
    return out;
}
std::string HF_Tokenizer::stream_flush(std::string* pending) {
    std::string out = utf8_lossy(*pending);
    pending->clear();
    return out;
}
int HF_Tokenizer::token_to_id(const std::string& token) const {
    auto it = tok2id_.find(token);
    return it == tok2id_.end() ? -1 : it->second;
}
std::string HF_Tokenizer::id_to_token(int id) const { return id >= 0 && id < (int)id2tok_.size() ? id2tok_[id] : std::string(); }

// ------------------------------------------------------------------------------------------------ loading
static void need(bool cond, const std::string& what) {
    if (!cond) throw std::runtime_error("tokenizer.json: " + what);
}
static std::string type_of(const JSON& j) { return j.is_object() && j.contains("type") ? j.at("type").as_string() : std::string(); }

std::shared_ptr<HF_Tokenizer> HF_Tokenizer::FromJSONText(const std::string& text, const std::string& config_text, std::string* err) {
    try {
        const JSON j = JSON::parse(text);
        need(j.is_object() && j.contains("model"), "no \"model\"");
        auto tk = std::shared_ptr<HF_Tokenizer>(new HF_Tokenizer());
        // normalizer: none, NFC, or a Sequence of NFC
        if (const JSON* nz = j.find("normalizer")) {
            if (!nz->is_null()) {
                std::vector<const JSON*> items;
                if (type_of(*nz) == "Sequence")
                    for (const JSON& x : nz->at("normalizers").arr) items.push_back(&x);
                else
                    items.push_back(nz);
                for (const JSON* x : items) {
                    need(type_of(*x) == "NFC", "normalizer '" + type_of(*x) + "' is not built (none or NFC)");
                    tk->nfc_ = true;
                }
            }
        }
        // pre_tokenizer: Sequence[Split(pattern, Isolated), ByteLevel(use_regex = false)]
        const JSON* pt = j.find("pre_tokenizer");
        need(pt && type_of(*pt) == "Sequence" && pt->contains("pretokenizers"), "pre_tokenizer must be Sequence[Split, ByteLevel]");
        const auto& pts = pt->at("pretokenizers").arr;
        need(pts.size() == 2 && type_of(pts[0]) == "Split" && type_of(pts[1]) == "ByteLevel", "pre_tokenizer must be Sequence[Split, ByteLevel]");
        const JSON* rx = pts[0].path({"pattern", "Regex"});
        need(rx && rx->is_string(), "Split pattern must be a Regex");
        const std::string& pat = rx->str;
        if (pat == kPatternQwen)
            tk->max_digits_ = 1;
        else if (pat == kPatternLlama3)
            tk->max_digits_ = 3;
        else
            need(false, "Split pattern is not one of the two this scanner implements (Qwen2/3, Llama-3): " + rx->str);
        need(pts[0].contains("behavior") && pts[0].at("behavior").as_string() == "Isolated", "Split behavior must be Isolated");
        need(!pts[0].contains("invert") || !pts[0].at("invert").as_bool(false), "Split invert must be false");
        need(!pts[1].contains("add_prefix_space") || !pts[1].at("add_prefix_space").as_bool(false), "ByteLevel add_prefix_space must be false");
        need(pts[1].contains("use_regex") && !pts[1].at("use_regex").as_bool(true), "ByteLevel use_regex must be false");
        if (const JSON* dc = j.find("decoder")) need(dc->is_null() || type_of(*dc) == "ByteLevel", "decoder must be ByteLevel");
        // model: BPE
        const JSON& m = j.at("model");
        need(type_of(m) == "BPE" || (!m.contains("type") && m.contains("merges")), "model must be BPE");
        need(!m.contains("byte_fallback") || !m.at("byte_fallback").as_bool(false), "BPE byte_fallback must be false");
        for (const char* k : {"continuing_subword_prefix", "end_of_word_suffix"})
            if (const JSON* v = m.find(k)) need(v->is_null() || (v->is_string() && v->str.empty()), std::string("BPE ") + k + " must be empty");
        if (const JSON* v = m.find("ignore_merges")) tk->ignore_merges_ = v->as_bool(false);
        const JSON& vocab = m.at("vocab");
        need(vocab.is_object(), "BPE vocab must be an object");
        int max_id = -1;
        tk->tok2id_.reserve(vocab.obj.size() * 2);
        for (auto& kv : vocab.obj) {
            const int id = kv.second.as_int(-1);
            need(id >= 0, "vocab id of '" + kv.first + "'");
            tk->tok2id_[kv.first] = id;
            max_id = std::max(max_id, id);
        }
        if (const JSON* at = j.find("added_tokens"))
            for (const JSON& a : at->arr) {
                AddedToken t;
                t.content    = a.at("content").as_string();
                t.id         = a.at("id").as_int(-1);
                t.special    = a.contains("special") && a.at("special").as_bool(false);
                t.normalized = a.contains("normalized") && a.at("normalized").as_bool(false);
                need(t.id >= 0 && !t.content.empty(), "added token without id / content");
                for (const char* k : {"lstrip", "rstrip", "single_word"})
                    need(!a.contains(k) || !a.at(k).as_bool(false), "added token '" + t.content + "': " + k + " is not built");
                tk->added_.push_back(t);
                max_id = std::max(max_id, t.id);
            }
        tk->id2tok_.assign((size_t)max_id + 1, std::string());
        tk->special_.assign((size_t)max_id + 1, 0);
        for (auto& kv : tk->tok2id_) tk->id2tok_[kv.second] = kv.first;
        for (const AddedToken& t : tk->added_) {
            tk->id2tok_[t.id]      = t.content;
            tk->tok2id_[t.content] = t.id;
            tk->special_[t.id]     = t.special;
        }
        const ByteChars& bc = byte_chars();
        for (int b = 0; b < 256; b++) {
            auto it = tk->tok2id_.find(bc.utf8[b]);
            need(it != tk->tok2id_.end(), "byte-level vocabulary lacks the character of byte " + std::to_string(b));
            tk->byte_tok_[b] = it->second;
        }
        // merges: "a b" strings (older files) or ["a", "b"] pairs; rank = position
        if (const JSON* mg = m.find("merges")) {
            int rank = 0;
            tk->merges_.reserve(mg->arr.size() * 2);
            for (const JSON& e : mg->arr) {
                std::string a, b;
                if (e.is_string()) {
                    const size_t sp = e.str.find(' ');
                    need(sp != std::string::npos, "merge '" + e.str + "'");
                    a = e.str.substr(0, sp), b = e.str.substr(sp + 1);
                } else {
                    need(e.is_array() && e.arr.size() == 2, "merge entry must be \"a b\" or [a, b]");
                    a = e.arr[0].as_string(), b = e.arr[1].as_string();
                }
                auto ia = tk->tok2id_.find(a), ib = tk->tok2id_.find(b), iab = tk->tok2id_.find(a + b);
                need(ia != tk->tok2id_.end() && ib != tk->tok2id_.end() && iab != tk->tok2id_.end(), "merge (" + a + ", " + b + ") names tokens outside the vocabulary");
                tk->merges_.emplace(((uint64_t)(uint32_t)ia->second << 32) | (uint32_t)ib->second, std::make_pair(rank, iab->second));
                rank++;
            }
        }
        // eos / bos / pad: tokenizer_config.json when given, else the family's usual names
        auto named = [&](const JSON& cfg, const char* key) -> int {
            const JSON* v = cfg.find(key);
            if (!v) return -1;
            const std::string s = v->is_string() ? v->str : (v->is_object() && v->contains("content")) ? v->at("content").as_string() : std::string();
            return s.empty() ? -1 : tk->token_to_id(s);
        };
        if (!config_text.empty()) {
            const JSON cfg = JSON::parse(config_text);
            tk->eos_ = named(cfg, "eos_token"), tk->bos_ = named(cfg, "bos_token"), tk->pad_ = named(cfg, "pad_token");
        }
        for (const char* s : {"<|im_end|>", "<|eot_id|>", "<|endoftext|>", "</s>"})
            if (tk->eos_ < 0) tk->eos_ = tk->token_to_id(s);
        for (const char* s : {"<|endoftext|>", "<pad>"})
            if (tk->pad_ < 0) tk->pad_ = tk->token_to_id(s);
        return tk;
    } catch (const std::exception& e) {
        if (err) *err = e.what();
        return nullptr;
    }
}
std::shared_ptr<HF_Tokenizer> HF_Tokenizer::FromPath(const std::string& path, std::string* err) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) {
        if (err) *err = "no such file or directory: '" + path + "'";
        return nullptr;
    }
    std::string file = path, dir;
    if (S_ISDIR(st.st_mode))
        dir = path, file = path + "/tokenizer.json";
    else
        dir = path.find('/') == std::string::npos ? "." : path.substr(0, path.rfind('/'));
    bool ok;
    const std::string text = read_file(file, &ok);
    if (!ok) {
        if (err) *err = "cannot open '" + file + "'";
        return nullptr;
    }
    bool has_cfg;
    const std::string cfg = read_file(dir + "/tokenizer_config.json", &has_cfg);
    return FromJSONText(text, has_cfg ? cfg : std::string(), err);
}

// ------------------------------------------------------------------------------------------------ ChatML
std::string ChatMLPrompt(const std::string& system, const std::string& user, bool enable_thinking) {
    std::string s;
    if (!system.empty()) s += "<|im_start|>system\n" + system + "<|im_end|>\n";
    s += "<|im_start|>user\n" + user + "<|im_end|>\n<|im_start|>assistant\n";
    if (!enable_thinking) s += "<think>\n\n</think>\n\n";
    return s;
}
std::string ChatMLRender(const std::vector<std::pair<std::string, std::string>>& lines, bool enable_thinking) {
    std::string result;
    for (const auto& line : lines) {
        result += "<|im_start|>" + line.first + "\n";
        if (line.first == "assistant") result += enable_thinking ? "\n\n" : "<think>\n\n</think>\n\n";
        result += line.second + "<|im_end|>\n";
    }
    return result;
}

}  // namespace koifish
