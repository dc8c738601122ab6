/*
 * kf_tokenizer.h -- C ABI of the text side of the chat loop: what a Koifish maintainer binds instead of
 * AutoTokenizer::from_pretrained / HF_Tokenizer::encode / decode / T2STR / eos_token_id (reference src/TokenSet/HF_Tokenizer.cpp:1723-1830)
 * and CHAT_SAMPLER::InitPrefillTemplate / toChatML (reference src/Utils/CLI_params.cpp:1990-2031).  Host only: no device, no CUDA.
 * Strings are UTF-8; every returned string is malloc'ed and freed with kf_string_free (kf_model.h).  Status codes are kf_device.h's.
 */
#ifndef KF_TOKENIZER_H
#define KF_TOKENIZER_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kf_tokenizer kf_tokenizer;

/* path: a tokenizer.json or a directory holding one (tokenizer_config.json next to it names eos / bos / pad).  The pipeline built is the
 * Qwen2 / Qwen3 (and Llama-3) one: literal added tokens -> NFC -> Split(GPT-4-style pattern) -> ByteLevel -> BPE; any other tokenizer.json
 * is refused with a message in *err_out. */
int kf_tokenizer_load(const char* path, kf_tokenizer** out, char** err_out);
/* the same from memory; config_json (tokenizer_config.json) may be NULL */
int kf_tokenizer_from_json(const char* tokenizer_json, const char* config_json, kf_tokenizer** out, char** err_out);
int kf_tokenizer_destroy(kf_tokenizer* t);
/* HF_Tokenizer::encode(text, add_special_tokens = false): ids_out holds `capacity` ids; *n_out receives the number the text encodes to
 * (KF_ERR_BAD_ARG with *n_out set when capacity is too small; ids_out may be NULL to size the buffer).  Invalid UTF-8 -> KF_ERR_BAD_ARG. */
int kf_tokenizer_encode(const kf_tokenizer* t, const char* text, size_t text_bytes, int32_t* ids_out, size_t capacity, size_t* n_out);
/* HF_Tokenizer::decode(ids, skip_special_tokens) */
int kf_tokenizer_decode(const kf_tokenizer* t, const int32_t* ids, size_t n, int skip_special_tokens, char** text_out);
/* Piece-by-piece printing (Fish::Chat prints tokenizer->T2STR(token) per step, GoPT.cpp:1203-1230): byte-level BPE splits rare characters and
 * emoji over several tokens, so T2STR of each alone prints U+FFFD pieces.  push returns the text that is certain after one more token and holds
 * back an unfinished multi-byte character; flush returns what is left (end of the answer).  The concatenation of every push and the flush equals
 * kf_tokenizer_decode of the whole sequence.  Returned strings are NUL-terminated (a decoded U+0000 ends them). */
typedef struct kf_decode_stream kf_decode_stream;
int kf_decode_stream_create(kf_decode_stream** out);
int kf_decode_stream_destroy(kf_decode_stream* s);
int kf_decode_stream_push(const kf_tokenizer* t, kf_decode_stream* s, int id, int skip_special_tokens, char** text_out);
int kf_decode_stream_flush(kf_decode_stream* s, char** text_out);
int kf_tokenizer_token_to_id(const kf_tokenizer* t, const char* token); /* -1 when absent */
int kf_tokenizer_id_to_token(const kf_tokenizer* t, int id, char** token_out);
int kf_tokenizer_vocab_size(const kf_tokenizer* t); /* largest id + 1, added tokens included */
int kf_tokenizer_eos_id(const kf_tokenizer* t);     /* tokenizer->S.eos of Fish::Chat (GoPT.cpp:1171): -1 when the file names none */
int kf_tokenizer_bos_id(const kf_tokenizer* t);
int kf_tokenizer_pad_id(const kf_tokenizer* t);
int kf_tokenizer_is_special(const kf_tokenizer* t, int id);

/* stages, for parity tests: Unicode NFC of a string; the pre-tokenisation pieces of a string joined by '\n'-free separators is not possible
 * in general, so pieces come back as a JSON array of strings */
int kf_text_nfc(const char* text, size_t text_bytes, char** out);
int kf_tokenizer_pre_tokenize(const kf_tokenizer* t, const char* text, size_t text_bytes, char** json_array_out);

/* CHAT_SAMPLER::InitPrefillTemplate: "<|im_start|>system\n%s<|im_end|>\n" (when system is non-empty) "<|im_start|>user\n%s<|im_end|>\n
 * <|im_start|>assistant\n" (+ "<think>\n\n</think>\n\n" when enable_thinking == 0) */
int kf_chatml_prompt(const char* system_or_null, const char* user, int enable_thinking, char** out);
/* CHAT_SAMPLER::toChatML over n (role, content) lines */
int kf_chatml_render(const char* const* roles, const char* const* contents, int n, int enable_thinking, char** out);

#ifdef __cplusplus
}
#endif
#endif /* KF_TOKENIZER_H */
