// QWen3.hpp -- host-side mirror of the reference's operator surface for Qwen3 decode / prefill:
//   MODEL_CARD / CLI_params keys   (src/CLI_params.hpp:263-385, src/Utils/CLI_params.cpp:1480-1545, 2224-2300)
//   neurons  SLP, LayerNormal, ROPE, SelfAttention, FFN, TokenEmbed, Head4Token (src/Manifold/Neuron.hpp:363-802) with the
//            cuInfer / Forw / cuFlow entry points of src/Device/CUDA/QKV.cu:617-706, NeuronFuse.cu:176-207, 305-381, 615-656,
//            842-862, T.cu:569-573, kernel/rope.cu:645-672
//   KVCache  (src/Utils/Cache.hpp:22-53, Cache.cpp:14-60)
//   QWen3 / Fish  (src/Transformer/QWen.cpp:16-145, src/Manifold/Fish.cpp:13-95, GoPT.cpp:1111-1235 Chat loop)
// plus what the reference does not have: tensor-parallel sharding (one process per GPU) and CUDA-graph replay of a token.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../Tensor/GTensor.hpp"

namespace koifish {

// ---- config -------------------------------------------------------------------------------------------------------------
struct MODEL_CARD {
    std::string arch = "QWEN3";
    int n_layers = 0, n_embd = 0, n_ff = 0, n_head = 0, n_head_kv = 0, head_dim = 128;
    int vocab = 151936;  // pad_vocab_size, CLI_params.hpp:327
    int max_pos_embeddings = 32768, n_ctx = 1024;
    int max_seq_len  = 1024;  // chat_sampler.seq_len default, CLI_params.hpp:697 ; JSON gpt.max_seq_len
    float rope_theta = 10000.f;  // random-init default (Neuron.cpp:612-620); HF card: config.json rope_theta
    float norm_rms_eps = 1e-6f;  // QWen.cpp:16-58
    bool tie_word_embeddings = false;
    bool isQKNormal = true, isSeparateQKV = true;  // QWen3 flags (QWen.cpp:16-58)
    JSON jQuant;                                    // the "quantizer" block
    int seed = 42;
    float init_sigma = 0.02f, norm_sigma = 0.0f;  // huTensor.cu:204 ; norms FIX_1
    int max_batch = 1;                            // independent sequences (batched decode)
    int max_prefill = 64;                         // tokens per prefill panel (gpt.max_prefill): sizes the activation buffers
    int awq_repack  = 0;                          // gpt.awq_repack: vendor AWQ tensors re-laid-out at load into PackedQ 4-bit storage (fast kernels)

    // accepts a Koifish JSON (cases/qwen3/*.json layout) or an HF config.json, optionally wrapped as {"hf_config": {...}}
    static MODEL_CARD FromJSON(const JSON& j);
    int q_dim() const { return n_head * head_dim; }
    int kv_dim() const { return n_head_kv * head_dim; }
};

struct Fish;
bool ShardPlan(const MODEL_CARD& c, const std::string& name, int rank, int world, int* rows_g, int* cols_g, int* rows_l, int* cols_l, int* row0,
               int* col0);

size_t AwqRepackBytes(int OCl, int ICl);
void AwqRepackWindow(const void* qweight, const void* qzeros, const void* scales, int IC, int OC, int r0, int OCl, int c0, int ICl, uint8_t* out_blob);
void AwqShardWindow(const void* qweight, const void* qzeros, const void* scales, int IC, int OC, int r0, int OCl, int c0, int ICl, uint8_t* out_blob);

// ---- neurons --------------------------------------------------------------------------------------------------------------
struct GeNeuron {
    std::string name;
    Fish* hFish = nullptr;
    virtual ~GeNeuron() {}
};
// SLP: linear neuron, y = x W^T (src/Manifold/Neuron.hpp:404; SLP::Forw NeuronFuse.cu:305-381)
struct SLP : GeNeuron {
    hGTensor w;
    int nIn = 0, nOut = 0;
    bool Empty() const { return !w; }
    // rhs[M][nOut] = lhs[M][nIn] . W^T ; epilogue: KF_EPI_*; returns KF status
    int Forw(void* rhs, const void* lhs, int M, int epilogue = KF_EPI_NONE, const void* residual = nullptr);
};
struct LayerNormal : GeNeuron {
    hGTensor w;
    float rms_eps = 1e-6f;
    int cuFlow(void* out, const void* inp, int rows);  // chat branch -> CU_rms_infer (T.cu:569-573)
};
struct SelfAttention;
struct ROPE : GeNeuron {
    hGTensor q_norm, k_norm;  // hnQ / hnK
    void* table = nullptr;    // (cos, sin)[max_seq][hd/2]
    float theta = 10000.f;
    int cuInfer(SelfAttention* hQKV, int M);  // rope.cu:645-672
};
struct SelfAttention : GeNeuron {
    int layid = 0;  // 1-based like the reference (QKV.cu:630 uses layid - 1)
    LayerNormal norm;
    SLP Q, K, V, proj_cat;
    ROPE rope;
    int n_head = 0, n_head_kv = 0, head_dim = 0;  // LOCAL (per tensor-parallel rank) head counts
    int cuInfer(void* inpL /* x in/out [M][E] */, int M);  // QKV.cu:617-706
};
struct FFN : GeNeuron {
    int layid = 0;
    LayerNormal norm;
    SLP gate, up, down;
    int latent = 0;  // LOCAL ffn width
    int cuInfer(void* inpL, int M);  // NeuronFuse.cu:615-656
};
struct TokenEmbed : GeNeuron {
    hGTensor w;
    int cuInfer(void* out, int M);  // NeuronFuse.cu:176-207
};
struct Head4Token : GeNeuron {
    LayerNormal norm;  // the final model.norm (a separate LayerNormal neuron in the reference graph)
    SLP proj;
    int cuInfer_1(void* logits, const void* inp, int M);  // NeuronFuse.cu:842-862
};
// KVCache: two bf16 tensors [nLayer][max_batch][max_seq][kv_dim_local] (Cache.cpp:14-27 has no batch dimension)
struct KVCache {
    void *key = nullptr, *value = nullptr;
    int n_layer = 0, max_batch = 1, max_seq = 0, kv_dim = 0;
    enum CTYPE { KV_KEY, KV_VAL };
    void* Get(CTYPE t, int layer, int pos = 0, int seq = 0) const;  // Cache.cpp:43-58
    size_t seq_stride() const { return (size_t)max_seq * kv_dim; }
    size_t bytes() const { return (size_t)2 * n_layer * max_batch * max_seq * kv_dim * 2; }
};

// ---- the model ------------------------------------------------------------------------------------------------------------
struct Fish {
    kf_ctx* ctx = nullptr;
    MODEL_CARD config;
    int tp_rank = 0, tp_world = 1;
    TokenEmbed embed;
    std::vector<std::unique_ptr<SelfAttention>> attn;
    std::vector<std::unique_ptr<FFN>> ffn;
    Head4Token cls;
    KVCache cache;
    std::map<std::string, hGTensor> tensors;  // by HF name (NN2NAME, QWen.cpp:61-145)
    std::map<std::string, int> tensor_ids;    // synthetic-weight seed ids (shared definition with the test oracle)

    // activations (views into one scratch allocation, like gBUFF; huTensor.cu:922-1003)
    int max_tokens = 0;
    void *x = nullptr, *xb = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *hb = nullptr, *logits = nullptr;
    float* part_f32 = nullptr;
    int32_t *d_tokens = nullptr, *d_pos = nullptr, *d_next = nullptr;
    // CHAT_SAMPLER (CLI_params.hpp:663-719): temperature 0 = greedy; d_rng = one xorshift64* state per sequence row
    float samp_temperature = 0.f, samp_top_p = 0.95f;
    int samp_top_k = 50, samp_selection = 0;
    uint64_t* d_rng = nullptr;
    int SetSampler(float temperature, int top_k, float top_p, uint64_t seed, int selection);
    int PickNext(int rows);  // d_next[r] = argmax or a sample of logits row r
    int32_t* h_stage  = nullptr;  // pinned staging: tokens | pos | next
    uint16_t* h_logits = nullptr; // pinned
    int seq_mode = 0;             // 0: the M tokens of a forward are one sequence (prefill) ; 1: M independent sequences
    int gqa_min_batch = 0;        // > 0: batched decode from this many sequences uses kf_attn_decode_gqa (env KF_GQA_MIN_BATCH, sweeps); 0: rule in cuInfer
    int logit_rows = 0;           // rows of the logits buffers (max(64, max_batch)); bigger panels report the last token only
    bool panel_consecutive = false;  // prefill panel at positions pos[0] .. pos[0] + M - 1 -> tensor-core flash attention
    bool last_only = false;       // mode 2: logits / argmax of the last token of the panel only
    bool tp_fuse = false;         // this forward runs the exchange as the epilogue of O / down (tensor parallel, <= 8 tokens, peer buffers)
    int attn_hint = 0;            // upper bound of the positions of the current forward (power-of-two bucket - 1): sizes the attention split
    int CtxBucket() const;        // log2 of that bucket; part of the graph key, so graphs are re-captured as the context grows
    std::map<int, kf_graph*> graphs;  // per (M, mode) replayable token graphs
    std::set<int> warm;               // signatures that already ran eagerly once (workspaces sized) -> next call captures
    bool use_graphs = true;
    void* rope_table_shared = nullptr;
    int staged_pos_max = 0;
    int staged_M       = 0;      // tokens / positions staged on the device by the last Forward() (0: nothing staged yet)
    uint64_t graph_gen = 0;      // kf_scratch_generation() the captured graphs were recorded under
    void SyncGraphGeneration();  // drop the graphs when a context scratch buffer they point to has been re-allocated since
    std::string error;
    size_t weight_bytes = 0;
    bool weights_dirty  = true;  // weight_bytes must be recounted (set by ResetGraphs, which every tensor-changing path calls)

    Fish(kf_ctx* ctx, const MODEL_CARD& card, int tp_rank, int tp_world);
    ~Fish();
    int Build();                 // allocate tensors + buffers (Fish::MakeInstance -> Build, Fish.cpp:13-95)
    int InitParamRandom();       // huTensor::InitParam random path + quantise at load
    int SetTensor(const std::string& hf_name, const void* host_bf16, int rows, int cols);  // SERIALIZE path: full (unsharded) tensor
    // vendor AWQ arrays of one linear (full shape), cut to this rank's window and kept in the vendor layout (GeQuant::ExTensor, GeQuant.cpp:144-200)
    int SetTensorAWQ(const std::string& hf_name, const void* qweight_i32, const void* qzeros_i32, const void* scales_f16, int in_features, int out_features);
    bool all_resident = false;  // every tensor has device data (checked once by Forward)
    hGTensor GetTensor(const std::string& hf_name) const;
    // one forward over M tokens already staged in d_tokens / d_pos (ForwardOnRLS, gLLM.cpp:755-769); logits for all M rows
    int ForwardOnRLS(int M, bool want_logits);
    // public step: host tokens/pos in, logits (optional) + greedy next tokens (optional) out.  H2D/D2H inside.
    int Forward(const int32_t* tokens, const int32_t* pos, int M, int seq_mode, uint16_t* logits_out, int32_t* next_out);
    // the resident tensors exactly as they sit in HBM (data || gama per tensor): SerialGamaData, reference huTensor.cu:413-458
    int SaveBlobs(const std::string& path);
    int LoadBlobs(const std::string& path);
    // the reference's own container (fish.kun, CKP_KOIFISH): the same payloads behind a safetensors header with szData / szGama per tensor
    int SaveKun(const std::string& path, const std::string& config_json);
    int LoadKun(const std::string& path, int* n_loaded, int* n_skipped);
    // Fish::Chat's generation loop (GoPT.cpp:1111-1235): prefill, then sample / stop on eos or a full window / feed back.  stop_reason: 1 eos,
    // 2 max_new tokens produced, 3 context window full
    int Generate(const int32_t* prompt, int n_prompt, int pos0, int max_new, int eos_id, int32_t* out, int* n_out, int* stop_reason);
    // device-resident greedy loop: n_steps graph replays feeding argmax back as the next token (no host round trip)
    int DecodeLoop(int n_steps, int M);
    int UseGraph(int M, bool want_logits);
    void ResetGraphs();
    int AllocTensor(const std::string& name, int rows, int cols, int id, hGTensor& out);
};

}  // namespace koifish
