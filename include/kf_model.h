/*
 * kf_model.h -- C ABI of the host-side Qwen3 runtime (koifish_b200/csrc/Transformer, csrc/Tensor), i.e. what a Koifish
 * maintainer binds instead of Fish::MakeInstance / Fish::Chat / Fish::ForwardOnRLS (reference src/Manifold/Fish.cpp:13-95,
 * src/Manifold/GoPT.cpp:1111-1235, src/Manifold/gLLM.cpp:706-787).  Plain pointers and sizes only; every call returns a
 * kf_device.h status code and never exits.  Token ids / positions / logits cross this boundary in HOST memory; the
 * host<->device copies happen inside the call.
 */
#ifndef KF_MODEL_H
#define KF_MODEL_H
#include "kf_device.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kf_model kf_model;

typedef struct kf_model_info {
    int n_layers, n_embd, n_ff, n_head, n_head_kv, head_dim, vocab, max_seq_len, max_batch, max_tokens;
    int tp_rank, tp_world;
    int tie_word_embeddings;
    float rope_theta, norm_rms_eps;
    uint64_t weight_bytes;   /* packed weights + gama resident on THIS rank */
    uint64_t kv_bytes;       /* KV cache bytes on this rank */
    uint64_t block_weight_bytes_per_layer; /* packed+gama bytes of one transformer block on this rank */
    uint64_t head_weight_bytes;            /* lm_head bytes read per token on this rank */
} kf_model_info;

/* config_json: a Koifish JSON config (keys model.arch, model.parameter.{Layer, transformer.{Ctx,Embed,Ffn,Head,KVHead,head_dim},
 * tie_word_embeddings, max_pos_embeddings}, quantizer.{group_size, <name-substring>:{quant_method,bits,group_size}}, seed,
 * gpt.max_seq_len -- reference cases/qwen3/qwen3_596M_q4.json, src/Utils/CLI_params.cpp:1480-1545, src/Tensor/GeQuant.cpp:1186-1285)
 * or an HF config.json (src/Utils/CLI_params.cpp:2224-2300).  Extensions: model.parameter.{vocab_size, rope_theta},
 * gpt.max_batch, init.{sigma, norm_sigma}.  On error *err_out (if given) receives a malloc'ed message (free with kf_string_free). */
int kf_model_create(kf_ctx* ctx, const char* config_json, int tp_rank, int tp_world, kf_model** out, char** err_out);
int kf_model_destroy(kf_model* m);
const char* kf_model_error(kf_model* m);
void kf_string_free(char* s);
int kf_model_info_get(kf_model* m, kf_model_info* out);

/* huTensor::InitParam random path (reference src/Device/CUDA/huTensor.cu:157-231): synthetic N(0, sigma^2)-like weights from the
 * framework's counter-based generator, quantised at load per the quantizer card (GeQuant::LowBit_worker, GeQuant.cpp:830-905) */
int kf_model_init_random(kf_model* m);
/* SERIALIZE path: hand over one FULL (unsharded) bf16 tensor by its HF name (NN2NAME, src/Transformer/QWen.cpp:61-145);
 * it is sharded for this rank and quantised at load */
int kf_model_set_tensor(kf_model* m, const char* hf_name, const void* host_bf16, int rows, int cols);
/* One linear of a vendor-quantised (AWQ) checkpoint by the HF name of its weight ("....q_proj.weight"), FULL shape, in the arrays the
 * checkpoint stores (GeQuant::ExTensor src/Tensor/GeQuant.cpp:144-200, GTensor::LoadParam src/Manifold/Serialize.cpp:145-230, unpack
 * CU_Q42X_awq src/Device/CUDA/kernel/quantizer.cu:132-156): qweight int32 [in][out / 8] (nibbles in AWQ_REVERSE_ORDER), qzeros int32
 * [in / 128][out / 8], scales fp16 [in / 128][out].  The quantizer card must select "quant_method": "awq" for the tensor (an HF config's
 * "quantization_config" does: QUANT_CARD::Vendor2JSONx, src/Utils/CLI_params.cpp:240-262).  The arrays are cut to the rank's window and
 * stay in the vendor layout on the device (type KF_T_AWQ4), read as CU_Q42X_awq reads them (bit-equal weights; a functional, untuned matmul).
 * With "gpt": {"awq_repack": 1} in the config they are instead re-laid-out at load into the library's own 4-bit storage (type KF_T_Q4: the
 * same codes in PackedQ words over [out][in], step = bf16(scale), zero = bf16(zero_point * scale)) so that the tuned decode / tensor-core
 * kernels run on them; the weights then differ from CU_Q42X_awq's by the bf16 rounding of step and zero (a few 1e-3 of the group's range). */
int kf_model_set_tensor_awq(kf_model* m, const char* hf_name, const void* qweight_i32, const void* qzeros_i32, const void* scales_f16,
                            int in_features, int out_features);
/* descriptor of the device-resident (possibly packed) tensor: for parity tests (GetDataX equivalent via kf_dequant) */
int kf_model_tensor_desc(kf_model* m, const char* hf_name, kf_tensor_desc* out);
int kf_model_tensor_count(kf_model* m);
const char* kf_model_tensor_name(kf_model* m, int index);
void* kf_model_kcache(kf_model* m, int layer);
void* kf_model_vcache(kf_model* m, int layer);

/* One forward over M tokens (Fish::ForwardOnRLS; one call per token in the reference's Chat loop, GoPT.cpp:1139-1146).
 *   seq_mode 0: the M tokens are consecutive positions of ONE sequence (prefill panel; M <= max_tokens);
 *   seq_mode 1: M independent sequences, one token each (batched decode; M <= gpt.max_batch);
 *   seq_mode 2: as 0, but logits_host / next_host receive ONE row: the last token of the panel (long prompts: gpt.max_prefill
 *               sizes the panel, default 64; consecutive positions run the tensor-core flash attention kf_attn_prefill).
 * tokens_host / pos_host: int32[M].  logits_host (optional): bf16 [M][vocab] (modes 0 / 1: M <= max(64, gpt.max_batch)).
 * next_host (optional): int32[M] greedy argmax.
 * Synchronous: returns after the results are in host memory. */
int kf_model_forward(kf_model* m, const int32_t* tokens_host, const int32_t* pos_host, int M, int seq_mode, void* logits_host,
                     int32_t* next_host);
/* n_steps greedy decode steps entirely on the device (each step a CUDA-graph replay feeding its argmax back as the next token),
 * continuing from the tokens/positions of the last kf_model_forward.  Asynchronous; kf_ctx_sync() to wait. */
int kf_model_decode_loop(kf_model* m, int n_steps, int M);
/* The generation loop of Fish::Chat (reference src/Manifold/GoPT.cpp:1111-1235): the prompt (int32[n_prompt], host) is prefilled at positions
 * pos0 .. pos0 + n_prompt - 1 (pos0 > 0 continues a conversation whose earlier turns are in the KV cache), then tokens are drawn with the
 * model's sampler (kf_model_set_sampler; greedy by default) and fed back until one equals eos_id (not emitted; pass -1 for none),
 * max_new_tokens have been produced, or the context window (gpt.max_seq_len) is full.  out_ids holds max_new_tokens ids; *n_out receives the
 * count; *stop_reason_out (optional): 1 eos, 2 max_new_tokens, 3 context window full.  Synchronous. */
int kf_model_generate(kf_model* m, const int32_t* prompt_ids, int n_prompt, int pos0, int max_new_tokens, int eos_id, int32_t* out_ids, int* n_out,
                      int* stop_reason_out);
/* current device-side tokens / positions (after a decode loop) */
int kf_model_read_state(kf_model* m, int32_t* tokens_host, int32_t* pos_host, int M);
int kf_model_set_graphs(kf_model* m, int enable);
/* CHAT_SAMPLER (src/CLI_params.hpp:663-719): how kf_model_forward's next token and kf_model_decode_loop's feedback token are drawn.
 * temperature 0 (the default) = greedy.  Every sequence row starts from the same seed, as the reference's LogitsInfo::rng_state
 * (src/Manifold/GoPT.cpp:709).  See kf_sample for `selection`. */
int kf_model_set_sampler(kf_model* m, float temperature, int top_k, float top_p, uint64_t seed, int selection);
/* Save / load every resident tensor exactly as it sits in HBM (packed data || gama, the reference's per-tensor SerialGamaData payload,
 * src/Device/CUDA/huTensor.cu:413-458): loading skips the quantiser.  The file must come from a model built from the same config
 * (names, shapes, storage types and groups are checked); tensor-parallel ranks use one file per rank. */
/* HF checkpoints (Fish::LoadFolderOfST -> SAFETENSOR2Gensors -> GTensor::LoadParam, src/Manifold/Serialize.cpp:1010-1100, :145-230):
 * every tensor of `path_or_dir` ("model.safetensors", or a directory of *.safetensors shards) whose name the model knows is converted
 * to bf16 (BF16 / F16 / F32 sources), sharded for this rank and quantised per the quantizer card, as kf_model_set_tensor does.  Unknown
 * names are skipped and counted.  Vendor-quantised AWQ linears (<prefix>.qweight I32 / .qzeros I32 / .scales F16, possibly in different
 * shards) become <prefix>.weight in the AWQ layout as kf_model_set_tensor_awq does; each complete triple counts as one loaded tensor.
 * kf_safetensors_index: the header of one file as JSON text [{"name","dtype","shape","nbytes"}, ...] (host only; free with kf_string_free). */
int kf_model_load_safetensors(kf_model* m, const char* path_or_dir, int* n_loaded_out, int* n_skipped_out);
int kf_safetensors_index(const char* path, char** json_out, char** err_out);
/* one tensor of one file converted to bf16 exactly as the loader converts it (host only) */
int kf_safetensors_read_bf16(const char* path, const char* name, void* out_bf16_host, size_t capacity_elems, char** err_out);
int kf_model_save(kf_model* m, const char* path);
int kf_model_load(kf_model* m, const char* path);
/* The reference's own container, "fish.kun" (CKP_KOIFISH): a safetensors file whose header entries are {"dtype": K_FLOATS name ("Q<4>",
 * "TERNARY", "BINARY", "F8E5M2", "BF16(E8)" ...; src/g_float.hpp:127-151), "shape", "data_offsets", "loAB", "szGama", "szData"} (GTensor::jDesc,
 * src/Manifold/Serialize.cpp:61-100), whose payloads are the device blobs data || gama (GTensor::SerialGamaData, src/Device/CUDA/huTensor.cu:
 * 413-458) and whose "__koifish__config__" entry is the writer's JSON config as msgpack (K_SafeTensors::insertJS, src/Tensor/Safetensors.hpp:
 * 87-102).  kf_model_save_kun writes every resident tensor that way (AWQ tensors are refused: they have a checkpoint format of their own);
 * kf_model_load_kun reads such a file -- this library's or the reference's -- into a model built from a matching config (dtype, shape, szData
 * and szGama are checked against what the config selects; unknown names are skipped and counted; training-state files, whose payloads carry
 * optimizer moments, are refused).  Loading skips the quantiser, like the reference's Serial_Quant_MMAP when the stored type is the card's. */
int kf_model_save_kun(kf_model* m, const char* path);
int kf_model_load_kun(kf_model* m, const char* path, int* n_loaded_out, int* n_skipped_out);
/* host only: the header of a .kun file as JSON text [{"name","dtype","shape","szData","szGama","offset"}, ...]; its config entry decoded from
 * msgpack to JSON text ("" when absent); and a writer from host blobs (shapes: n x 2, second 0 for a vector; blobs[i] = szData + szGama bytes) */
int kf_kun_index(const char* path, char** json_out, char** err_out);
int kf_kun_config(const char* path, char** json_out, char** err_out);
int kf_kun_write(const char* path, const char* config_json, int n, const char* const* names, const char* const* dtypes, const int64_t* shapes,
                 const uint64_t* sz_data, const uint64_t* sz_gama, const void* const* blobs, char** err_out);

/* Host-only config logic (no device needed): parse a config and report the model dimensions, and which storage type the
 * quantizer block selects for a tensor name (QUANT_CARD::Init4Neuron, reference src/Tensor/GeQuant.cpp:1186-1285; MakeInstance
 * :23-81).  type_out: KF_T_* (KF_T_BF16 when the tensor is not quantised); mode_out: KF_Q_*. */
int kf_config_dims(const char* config_json, kf_model_info* out, char** err_out);
/* the "quantizer" block the config resolves to, as JSON text ("" when none): an HF config's "quantization_config" goes through the mapping of
 * QUANT_CARD::Vendor2JSONx (reference src/Utils/CLI_params.cpp:240-262) */
int kf_config_quantizer_json(const char* config_json, char** json_out, char** err_out);
/* the quantizer card QUANT_CARD::Init4Neuron fills for a tensor name, field by field (parity hook): out[8] = {selected, QUANT_MODE in the
 * reference's numbering (0 none, 1 RTN, 2 AWQ, 3 RTNf, 5 F8Ex; src/CLI_params.hpp:479-492), default_bits, T_group, yyang, isSymmetric,
 * isZeroPoint, isVendorQuant}; *errq_out = T_errQ */
int kf_config_quant_card(const char* config_json, const char* tensor_name, int* out, float* errq_out, char** err_out);
int kf_config_quant_of(const char* config_json, const char* tensor_name, int* type_out, int* group_out, int* mode_out, int* qbias_out,
                       char** err_out);
/* tensor-parallel shard plan: shape_out[6] = {rows_global, cols_global, rows_local, cols_local, row0, col0} of `tensor_name` on
 * rank `rank` of `world` (Q/K/V/gate/up split by output rows, O/down by input columns in whole quant groups, the rest replicated) */
int kf_config_shard_of(const char* config_json, const char* tensor_name, int rank, int world, int* shape_out, char** err_out);

/* the same plan applied to a vendor AWQ linear (host only): rank `rank`'s blob qweight || qzeros || scales -- exactly the bytes
 * kf_model_set_tensor_awq uploads -- cut from the FULL arrays.  *bytes_out = blob size; out_blob may be NULL to query it.  When the config
 * says gpt.awq_repack = 1 the blob is the re-laid-out one (PackedQ 4-bit data || gama of the window). */
int kf_config_awq_shard(const char* config_json, const char* tensor_name, int rank, int world, const void* qweight_i32, const void* qzeros_i32,
                        const void* scales_f16, void* out_blob, size_t capacity, size_t* bytes_out, char** err_out);

#ifdef __cplusplus
}
#endif
#endif /* KF_MODEL_H */
