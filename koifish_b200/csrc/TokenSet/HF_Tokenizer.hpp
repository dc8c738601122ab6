// HF_Tokenizer.hpp -- the text side of the chat loop: a reader of HF "tokenizer.json" files of the Qwen family and the ChatML
// prompt templates.  Mirrors the surface the reference's chat path uses:
//   HF_Tokenizer::encode / decode / T2STR / token_to_id / id_to_token / eos_token_id   (reference src/TokenSet/HF_Tokenizer.cpp:1723-1771)
//   AutoTokenizer::from_pretrained(dir) -> dir/tokenizer.json                           (:1813-1830)
//   CHAT_SAMPLER::InitPrefillTemplate / toChatML                                        (reference src/Utils/CLI_params.cpp:1990-2031)
// The reference file is a general port of the HF `tokenizers` crate (BPE, WordPiece, Unigram, seven pre-tokenizers, Oniguruma regex).  This one
// builds the single pipeline the Qwen3 / Qwen2.5 (and Llama-3-style) checkpoints of the hot path ship, and refuses anything else at load:
//   added tokens (literal, leftmost-longest)  ->  NFC  ->  Split(<the GPT-4-style pattern>, isolated)  ->  ByteLevel  ->  BPE (ranked merges)
//   decode: ByteLevel.
// The pattern is matched by a hand-written scanner over Unicode tables generated from unicodedata (unicode_tables.hpp); there is no regex engine.
// Host code only: no device, no CUDA.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace koifish {

struct AddedToken {
    std::string content;
    int id          = -1;
    bool special    = false;
    bool normalized = false;
};

class HF_Tokenizer {
   public:
    // `path`: a tokenizer.json, or a directory holding one (AutoTokenizer::from_pretrained); tokenizer_config.json / generation_config.json next
    // to it name the eos / bos / pad tokens when present.  nullptr + *err on anything the pipeline above does not cover.
    static std::shared_ptr<HF_Tokenizer> FromPath(const std::string& path, std::string* err);
    static std::shared_ptr<HF_Tokenizer> FromJSONText(const std::string& tokenizer_json, const std::string& tokenizer_config_json, std::string* err);

    // text must be valid UTF-8 (throws std::runtime_error otherwise)
    std::vector<int> encode(const std::string& text) const;
    std::string decode(const std::vector<int>& ids, bool skip_special_tokens) const;
    std::string T2STR(int id) const { return decode({id}, false); }  // the printable piece of one token (Fish::Chat prints these)
    // piece-by-piece printing without broken characters: push returns the text that is certain after this token, flush what is left
    std::string stream_push(std::string* pending, int id, bool skip_special_tokens) const;
    static std::string stream_flush(std::string* pending);
    std::string token_bytes(int id, bool skip_special_tokens) const;  // the raw bytes a token stands for
    int token_to_id(const std::string& token) const;                 // -1 when absent
    std::string id_to_token(int id) const;                           // "" when absent
    int vocab_size() const { return (int)id2tok_.size(); }           // largest id + 1 (model vocab + added tokens)
    int eos_token_id() const { return eos_; }
    int bos_token_id() const { return bos_; }
    int pad_token_id() const { return pad_; }
    bool is_special(int id) const { return id >= 0 && id < (int)special_.size() && special_[id]; }

    // the stages, exposed for tests
    static std::string NFC(const std::string& utf8);
    std::vector<std::string> pre_tokenize(const std::string& utf8) const;  // Split(pattern, isolated) pieces, before the byte-level mapping

   private:
    struct PairHash {
        size_t operator()(uint64_t k) const { return (size_t)(k * 0x9E3779B97F4A7C15ull >> 17); }
    };
    std::unordered_map<std::string, int> tok2id_;
    std::vector<std::string> id2tok_;
    std::vector<uint8_t> special_;
    std::unordered_map<uint64_t, std::pair<int, int>, PairHash> merges_;  // (left id << 32 | right id) -> (rank, merged id)
    std::vector<AddedToken> added_;
    int byte_tok_[256];  // token id of each single byte's byte-level character
    bool nfc_ = false, ignore_merges_ = false;
    int max_digits_ = 1;  // \p{N} (Qwen) or \p{N}{1,3} (Llama-3 / cl100k)
    int eos_ = -1, bos_ = -1, pad_ = -1;

    void bpe_word(const std::string& piece_utf8, std::vector<int>* out) const;
    void encode_plain(const std::string& utf8, std::vector<int>* out) const;  // text between added tokens (already normalised)
    void split_added(const std::string& text, bool normalized_pass, std::vector<std::pair<std::string, int>>* parts) const;
};

// CHAT_SAMPLER::InitPrefillTemplate (reference src/Utils/CLI_params.cpp:1990-2008): the single-turn prompt the reference's chat loop fills in.
// system may be empty (then the user-only template).  enable_thinking = false appends the empty think block so the model answers directly.
std::string ChatMLPrompt(const std::string& system, const std::string& user, bool enable_thinking);
// CHAT_SAMPLER::toChatML (:2010-2031): a whole conversation, one block per (role, content) line, exactly as the reference renders it
std::string ChatMLRender(const std::vector<std::pair<std::string, std::string>>& lines, bool enable_thinking);

}  // namespace koifish
