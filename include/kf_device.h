/*
 * kf_device.h -- the C ABI of koifish_b200's device layer (the drop-in boundary for Koifish's quantized-inference
 * hot path on B200 / sm_100a).
 *
 * The reference has no FFI layer: the path sits behind C++ neuron/tensor calls.  Each entry
 * point below names the reference interface it replaces (file:line relative to the reference tree).  Conventions:
 *   - plain pointers and sizes only; every pointer called "dev" is a CUDA device pointer, "host" a host pointer;
 *   - all bf16 tensors are raw uint16 bit patterns (floatX == floatGama == __nv_bfloat16, src/g_float.hpp:246-261);
 *   - every call returns an int status: 0 = KF_OK, negative = error (never exit(), unlike the reference's
 *     cudaCheck -> exit(KOIFISH_*), src/Device/CUDA/cuda_common.h:44-76);
 *   - kernels are launched on the context's stream; no call synchronises unless it says so;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns KF_ERR_NO_DEVICE.
 */
#ifndef KF_DEVICE_H
#define KF_DEVICE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference: exit codes in src/g_def_x.hpp:21-83, e.g. KOIFISH_QUANT_ERR -701) ---- */
enum {
    KF_OK              = 0,
    KF_ERR_NO_DEVICE   = -100,
    KF_ERR_CUDA        = -101,
    KF_ERR_BAD_ARG     = -102,
    KF_ERR_UNSUPPORTED = -103,
    KF_ERR_OOM         = -104,
    KF_ERR_NCCL        = -105,
    KF_ERR_QUANT       = -701,
};

/* ---- tensor storage types (typNUMBER, src/g_float.hpp:84-117; Bits2Type :177-192) ---- */
enum {
    KF_T_BF16     = 0, /* typNUMBER::BF16     16-bit, no quant                                   */
    KF_T_F8E5M2   = 1, /* typNUMBER::F8E5M2   8-bit = high byte of fp16 (packedN.cuh:80-96)      */
    KF_T_Q4       = 2, /* typNUMBER::Q4       4-bit codes in 128-bit words (PackedQ.hpp:99-183)  */
    KF_T_Q2       = 3, /* typNUMBER::Q2       2-bit RTN codes in 128-bit words                   */
    KF_T_SIGN     = 4, /* typNUMBER::T_SIGN   2-bit ternary {-1,0,1}+1 (yyang / bitnet)          */
    KF_T_BINARY   = 5, /* typNUMBER::T_BINARY 1-bit {0,1} (yyang)                                */
    KF_T_NF4      = 6, /* typNUMBER::Q4 under QUANT_MODE::RTNf ({"bits": 4} without a quant_method): NormalFloat4 codes as an MSB-first
                          nibble stream (BIT_SET_k, CLI_params.cpp:2177-2191), gama = [R_SCALE rows][C_SCALE cols][rows][16] bf16
                          per-row codebooks (GeQuant::_row_lut, GeQuant.cpp:696-732; CU_Q42X_NF4, quantizer.cu:612-654)            */
    KF_T_AWQ4     = 7, /* typNUMBER::Q4 under QUANT_MODE::AWQ: the vendor layout, STORED [in_features][out_features] (SLP::Forw uses
                          transA = 0 for it): data_dev = qweight int32 [cols][rows / 8] with AWQ_REVERSE_ORDER nibbles, zero_dev = qzeros
                          int32 [cols / 128][rows / 8], step_dev = scales fp16 [cols / 128][rows]; w = bf16(float(q - z) * scale)
                          (CU_Q42X_awq, quantizer.cu:132-156; CU_I2Q4_unpack, packedN.cuh:109-116).  rows = out_features, cols =
                          in_features, group = 128.  Device-level only: kf_dequant / kf_linear*; no quantiser (vendor checkpoints)   */
};

/* ---- quantisation modes (QUANT_CARD, src/CLI_params.hpp:509-554; GeQuant ctor src/Tensor/GeQuant.cpp:107-124) ---- */
enum {
    KF_Q_RTN_ASYM = 0,
    KF_Q_RTN_SYM  = 1,
    KF_Q_YYANG    = 2,
};

/* A (possibly quantised) weight W[rows = N_out][cols = K_in], row-major, as the reference stores it: ONE blob
 * data || gama (src/Tensor/GTensor.cpp:456-510, :1017):  gama = bf16 [R_SCALE rows][C_SCALE cols][ZERO nG][STEP nG],
 * nG = rows*cols/group.  gama may be NULL for BF16 / F8E5M2.  Mirrors what TASKA_quant passes to the reference's kernels
 * (src/Tensor/GeQuant.hpp:147-177). */
typedef struct kf_tensor_desc {
    const void* data_dev;
    const void* gama_dev;
    int rows, cols;
    int type;  /* KF_T_* */
    int group; /* T_group, 128 */
    int qbias; /* stored code = qid + qbias */
    /* optional: explicit ZERO / STEP arrays (bf16, one per group, row-major over this tensor's rows).  When NULL they are
     * located inside gama_dev by the blob layout above.  Used for row-slice views of a tensor (vocab-sharded tied lm_head). */
    const void* zero_dev;
    const void* step_dev;
} kf_tensor_desc;

typedef struct kf_ctx kf_ctx;

/* ---- context: replaces InitCUDA / main_stream / gBUFF scratch (src/Device/CUDA/QKV.cu:501-571, huTensor.cu:922-1003) ---- */
int kf_ctx_create(int device, void* cuda_stream /* cudaStream_t or NULL: the context creates its own */, kf_ctx** out);
int kf_ctx_destroy(kf_ctx* ctx);
int kf_ctx_sync(kf_ctx* ctx);
/* One process per GPU is the intended use.  A process that holds contexts on several devices must make the context's device current
 * (cudaSetDevice) before calling into it; kf_ctx_make_current does that.  The model runtime (kf_model.h) calls it on every forward. */
int kf_ctx_make_current(kf_ctx* ctx);
void* kf_ctx_stream(kf_ctx* ctx);
int kf_ctx_sm_count(kf_ctx* ctx);
const char* kf_status_string(int status);
const char* kf_last_error(kf_ctx* ctx);
/* number of kernels this library has launched on ctx since creation (bench.py's gpu_launches) */
uint64_t kf_launch_count(kf_ctx* ctx);
/* bumped whenever the context re-allocates one of its scratch buffers (split-K / attention workspaces, tensor-core staging): a CUDA graph
 * captured before the change still points at the freed buffer and must be re-captured.  The model runtime (kf_model.h) checks this before
 * every replay; callers that capture their own graphs with kf_graph_begin / kf_graph_end must do the same. */
uint64_t kf_scratch_generation(kf_ctx* ctx);
/* tuning knobs for sweeps: "gemv_splitk" (0 = heuristic), "gemv_variant", "gemv_exact" (0: factored dequant, NOT reference-exact),
 * "gemv_cluster" (single-token split-K merged inside a thread-block cluster), "attn_split", "attn_warps", "pdl" (programmatic
 * dependent launch), "deq_fma" (1, default: the dequant expression step*k - zero rounds ONCE in bf16, as the reference's kernel does when built
 * for sm_90+; 0: twice, as its -fmad=false / pre-sm_90 builds), "gemv_tma" (1, default: decode GEMVs of 4-bit weights take the persistent
 * TMA-fed stream-K kernel; 0: the round-1 cp.async kernel), "gemv_tma_occ" (CTAs per SM, 1 or 2), "gemv_tma_smem_kb" (shared-memory budget),
 * "gqa_min_ctx" (single-sequence decode: context length beyond which the kv-group tensor-core attention replaces the fused per-head
 * kernel; default 1024), "tc_min_m" (token count from which kf_linear* use the tcgen05 GEMM: -1 = measured per-type crossover, 0 = never, n = from n) */
int kf_ctx_set_int(kf_ctx* ctx, const char* key, int value);
int kf_ctx_get_int(kf_ctx* ctx, const char* key, int* value_out); /* gqa_min_ctx, tc_min_m, pdl, attn_split */

/* ---- device memory (huTensor::Alloc_1, src/Device/CUDA/huTensor.cu:70-103) ---- */
int kf_malloc(kf_ctx* ctx, size_t bytes, void** dev_out);
int kf_free(kf_ctx* ctx, void* dev);
int kf_memset(kf_ctx* ctx, void* dev, int value, size_t bytes);
int kf_h2d(kf_ctx* ctx, void* dev, const void* host, size_t bytes);  /* async on the stream (pinned host) or staged */
int kf_d2h(kf_ctx* ctx, void* host, const void* dev, size_t bytes);  /* async on the stream; call kf_ctx_sync before reading */
int kf_d2d(kf_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
int kf_host_alloc(size_t bytes, void** host_out); /* pinned */
int kf_host_free(void* host);

/* ---- CUDA graph capture of a launch sequence (no reference equivalent: it launches ~40 kernels/layer eagerly) ---- */
typedef struct kf_graph kf_graph;
int kf_graph_begin(kf_ctx* ctx);
int kf_graph_end(kf_ctx* ctx, kf_graph** out);
int kf_graph_launch(kf_ctx* ctx, kf_graph* g);
int kf_graph_destroy(kf_graph* g);

/* ---- synthetic weights: replaces CU_disti_normal in huTensor::InitParam (src/Device/CUDA/huTensor.cu:199-210) ---- */
int kf_fill_normal(kf_ctx* ctx, void* out_bf16_dev, size_t n, uint64_t seed, float sigma, float mean);
/* a [rows, cols] window at (row0, col0) of a virtual row-major matrix with ld_global columns: element (r, c) takes the
 * generator index (row0 + r) * ld_global + col0 + c, so a tensor-parallel shard equals the slice of the full tensor */
int kf_fill_normal_2d(kf_ctx* ctx, void* out_bf16_dev, int rows, int cols, size_t ld_global, size_t row0, size_t col0, uint64_t seed,
                      float sigma, float mean);

/* ---- quantise at load: GeQuant::LowBit_worker with flag 0x100 (source on GPU), src/Tensor/GeQuant.cpp:830-905;
 *      arithmetic of RTN_x :428-533 / YinYang :536-628 ; 8-bit: huTensor::ToF8Ex src/Device/CUDA/huTensor.cu:821-850 ---- */
size_t kf_quant_data_bytes(int rows, int cols, int type);
size_t kf_quant_gama_bytes(int rows, int cols, int type, int group);
int kf_quantize(kf_ctx* ctx, const void* w_bf16_dev, int rows, int cols, int type, int group, int mode, void* data_dev, void* gama_dev,
                int* qbias_out);

/* ---- GTensor::GetDataX (src/Device/CUDA/kernel/quantizer.cu:249-392): dequantise the whole weight to bf16 [rows, cols].
 *      Test hook only: the product never materialises dequantised weights. ---- */
int kf_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out_bf16_dev);

/* ---- TASKA_AxB::blasLt as used by SLP::Forw (src/Tensor/GTensor.hpp:703-741, src/Device/CUDA/NeuronFuse.cu:305-381):
 *      y[M][rows] = x[M][cols] . deq(W)^T, fp32 accumulate, bf16 out.  The reference dequantises W to a scratch and calls
 *      cuBLASLt; here unpack + dequant are fused into the matmul.
 *      epilogue flags: KF_EPI_RESIDUAL  y = RN(residual + RN_bf16(acc))   (replaces the following CU_add3, packedN.cuh:867-875)
 *      Few tokens take the HBM-bound skinny kernel (mma.sync GEMV), more the persistent tcgen05 / TMEM / TMA kernel; the crossover
 *      is per weight type (bf16: always tcgen05, everything else from 9 tokens; profiles/r01_tc_crossover.txt). ---- */
enum { KF_EPI_NONE = 0, KF_EPI_RESIDUAL = 1, KF_EPI_F32 = 4 /* y is float [M][rows], unrounded partial sums (tensor parallel) */ };
int kf_linear(kf_ctx* ctx, void* y_dev, const kf_tensor_desc* w, const void* x_dev, int M, int epilogue, const void* residual_dev);
/* TASKA_AxB in full (src/Tensor/GTensor.hpp:698-741): d = alpha * x . w^T + beta * d + bias (bias: one bf16 per output row, or NULL),
 * fp32 epilogue, one rounding -- what CU_mm_blasLt hands to cuBLASLt.  The inference path uses alpha 1, beta 0, no bias (= kf_linear). */
int kf_linear_axb(kf_ctx* ctx, void* d_dev, const kf_tensor_desc* w, const void* x_dev, int M, float alpha, float beta, const void* bias_dev);
/* up to 3 weights sharing x (Q/K/V: SelfAttention::cuInfer, src/Device/CUDA/QKV.cu:648-652) in one launch; y_dev[i] is [M][rows_i] */
int kf_linear_multi(kf_ctx* ctx, int n, void* const* y_dev, const kf_tensor_desc* w, const void* x_dev, int M);
/* FFN gate/up + CU_swiglu_v0 (src/Device/CUDA/NeuronFuse.cu:628-637, Activation.cu:86-93): y = silu(bf16(Wg x)) * bf16(Wu x) */
int kf_linear_swiglu(kf_ctx* ctx, void* y_dev, const kf_tensor_desc* w_gate, const kf_tensor_desc* w_up, const void* x_dev, int M);

/* LayerNormal::cuFlow (CU_rms_infer, layernorm.cuh:801-859) folded into the matmul(s) that consume its output, as
 * SelfAttention::cuInfer / FFN::cuInfer / Head4Token chain them (QKV.cu:640-652, NeuronFuse.cu:624-637): the normalised activations
 * are produced while x is staged on chip and never touch HBM.  Same arithmetic and rounding points as kf_rmsnorm followed by
 * kf_linear_multi (mode 0, n <= 3 outputs y[i] = [M][rows_i]) or kf_linear_swiglu (mode 2, w[0] gate, w[1] up -> y[0]). */
int kf_rmsnorm_linear(kf_ctx* ctx, int n, void* const* y_dev, const kf_tensor_desc* w, const void* x_dev, const void* norm_w_dev, float eps,
                      int M, int mode);

/* ---- LayerNormal::cuFlow chat branch -> CU_rms_infer (src/Device/CUDA/T.cu:569-573, kernel/layernorm.cuh:801-859) ---- */
int kf_rmsnorm(kf_ctx* ctx, void* out_dev, const void* x_dev, const void* w_dev, int rows, int dim, float eps);
/* ---- ROPE::cuInfer (src/Device/CUDA/kernel/rope.cu:645-672): per-head QK RMSNorm (layernorm.cuh:750-798), half-split RoPE
 *      (operator.cuh:735-772) on q and k, and the K/V rows written at `pos` of the cache (the reference aliases K.out/V.out
 *      onto the cache rows, src/Manifold/TGraph.cpp:198-208).  q is updated in place.  pos_dev: device int32[M] positions.
 *      rope_table_dev: float2 (cos, sin) [max_seq][head_dim/2] built by kf_rope_table. ---- */
int kf_rope_table(kf_ctx* ctx, void* table_dev, int max_seq, int head_dim, float theta);
int kf_qknorm_rope_kvappend(kf_ctx* ctx, void* q_dev, const void* k_dev, const void* v_dev, const void* qnorm_w_dev, const void* knorm_w_dev,
                            void* kcache_layer_dev, void* vcache_layer_dev, const void* rope_table_dev, const int32_t* pos_dev, int M,
                            int n_head, int n_kv, int head_dim, int max_seq, float eps, size_t seq_stride);
/* seq_stride (elements): token m uses the cache at base + m*seq_stride.  0 = all M tokens belong to ONE sequence (prefill panel:
 * token m attends to positions 0..pos[m]); max_seq*n_kv*head_dim = M independent sequences (batched decode). */
/* ---- attention_qk_kernel + CU_softmax_multihead + attention_v_kernel (src/Device/CUDA/kernel/operator.cuh:573-632, 252-277,
 *      650-668): GQA decode attention over a contiguous bf16 cache [max_seq][n_kv*hd] (KVCache, src/Utils/Cache.cpp:14-60),
 *      split-K over the sequence.  M query tokens; token m attends to positions 0..pos[m]. ---- */
int kf_attn_decode(kf_ctx* ctx, void* out_dev, const void* q_dev, const void* kcache_layer_dev, const void* vcache_layer_dev,
                   const int32_t* pos_dev, int M, int n_head, int n_kv, int head_dim, int max_seq, int max_pos_hint, size_t seq_stride);
/* ---- ROPE::cuInfer + the three attention kernels of SelfAttention::cuInfer (src/Device/CUDA/QKV.cu:660-674) in ONE launch for decode:
 *      QK-norm + RoPE of q and of the current k in registers, K / V appended at pos, split-K attention over the cache, slices merged
 *      by the last CTA.  Same arithmetic and rounding points as kf_qknorm_rope_kvappend followed by kf_attn_decode.  Precondition:
 *      M == 1, or the M tokens belong to M different sequences (seq_stride > 0).  q_dev is NOT updated (the reference's in-place
 *      normalised q is only ever consumed by this attention). ---- */
int kf_qkv_attention(kf_ctx* ctx, void* out_dev, const void* q_dev, const void* k_dev, const void* v_dev, const void* qnorm_w_dev,
                     const void* knorm_w_dev, void* kcache_layer_dev, void* vcache_layer_dev, const void* rope_table_dev,
                     const int32_t* pos_dev, int M, int n_head, int n_kv, int head_dim, int max_seq, float eps, size_t seq_stride,
                     int max_pos_hint);
/* ---- prefill attention: causal attention of a panel of M CONSECUTIVE tokens of one sequence (positions pos_dev[0] + m) over the
 *      cache rows [0, pos_dev[0] + M).  q_dev must already be normalised + rotated and the panel's K / V rows appended
 *      (kf_qknorm_rope_kvappend).  Replaces the reference's token-by-token prompt loop (src/Manifold/GoPT.cpp:1111-1235) through
 *      attention_qk / softmax / attention_v (src/Device/CUDA/kernel/operator.cuh:573-668); flash-attention on mma.sync tensor cores. ---- */
int kf_attn_prefill(kf_ctx* ctx, void* out_dev, const void* q_dev, const void* kcache_layer_dev, const void* vcache_layer_dev,
                    const int32_t* pos_dev, int M, int n_head, int n_kv, int head_dim, int max_seq);
/* ---- decode attention for many sequences per step: same contract as kf_attn_decode, but the query heads that share a kv head are
 *      processed together on the tensor cores, so every cached row is read once per kv head instead of once per query head
 *      (batched decode; replaces the same reference kernels, operator.cuh:573-668).  n_head / n_kv <= 16. ---- */
int kf_attn_decode_gqa(kf_ctx* ctx, void* out_dev, const void* q_dev, const void* kcache_layer_dev, const void* vcache_layer_dev,
                       const int32_t* pos_dev, int M, int n_head, int n_kv, int head_dim, int max_seq, int max_pos_hint, size_t seq_stride);
/* ---- CU_swiglu_v0 (Activation.cu:86-93), CU_add3 (packedN.cuh:867-875) as stand-alone ops ---- */
int kf_swiglu(kf_ctx* ctx, void* out_dev, const void* gate_dev, const void* up_dev, size_t n);
int kf_add(kf_ctx* ctx, void* out_dev, const void* a_dev, const void* b_dev, size_t n);
/* out = RN(residual + RN_bf16(sum_f32)): the residual add after a tensor-parallel all-reduce of fp32 partial sums */
int kf_residual_add_f32(kf_ctx* ctx, void* out_dev, const void* residual_dev, const float* sum_f32_dev, size_t n);
/* pos[m] += 1 ; used by the device-resident greedy decode loop */
int kf_advance_pos(kf_ctx* ctx, int32_t* pos_dev, int M);
/* ---- TokenEmbed::cuInfer (src/Device/CUDA/NeuronFuse.cu:176-207, kernel/embed.cuh:55-133): out[m] = deq(W[token[m]]) ---- */
int kf_embed(kf_ctx* ctx, void* out_dev, const kf_tensor_desc* w, const int32_t* tokens_dev, int M);
/* greedy sampler on device (the reference copies logits to the host, GoPT.cpp:614-630): out_token[m] = argmax(logits[m]) */
int kf_argmax(kf_ctx* ctx, int32_t* out_tokens_dev, const void* logits_bf16_dev, int M, int vocab);
/* Temperature / top-k / top-p sampling on the device, one token per logits row (GeneratOnPrompt::Sample, src/Manifold/GoPT.cpp:614-630 with
 * TopK :632-640, UpdateLogits :751-766, TopP :729-748, Qu_FlipCoin :768-786; the reference copies the logits to the host and samples there).
 * rng_state_dev: one 64-bit xorshift64* state per row (GoPT.cpp:594-600), advanced by the call.  temperature == 0 or top_k == 1: kf_argmax.
 * selection 0: the top_k largest logits, ties to the lower index; 1: what TOPK_heap::Select (GoPT.cpp:667-700) keeps as written (its heap
 * orders indices: {0 .. k-2} plus the first maximum of the rest).  top_k <= 1024. */
int kf_sample(kf_ctx* ctx, int32_t* out_tokens_dev, const void* logits_bf16_dev, int M, int vocab, float temperature, int top_k, float top_p,
              uint64_t* rng_state_dev, int selection);

/* ---- tensor parallel: no reference equivalent (multi_gpu.cuh is dead code).  One process per GPU;
 *      the unique id is exchanged by the caller (torch.distributed / MPI) ---- */
int kf_nccl_unique_id(void* id_out_128_bytes);
int kf_ctx_init_nccl(kf_ctx* ctx, const void* id_128_bytes, int rank, int world);
int kf_allreduce_bf16(kf_ctx* ctx, void* buf_dev, size_t count); /* in-place sum over ranks, on the stream */
/* all-gathered vocabulary slices [world][M][vl] (bf16) -> logits rows [M][world * vl] in one launch (vocab-sharded lm_head) */
int kf_relayout_wmv(kf_ctx* ctx, void* out_dev, const void* in_dev, int world, int M, int vl);
int kf_allreduce_f32(kf_ctx* ctx, float* buf_dev, size_t count);
int kf_allgather(kf_ctx* ctx, void* out_dev, const void* in_dev, size_t bytes_per_rank);
/* ---- the decode exchange over NVLink / NVSwitch PEER MEMORY, fused with the residual add (p2p.cu): one launch instead of
 *      ncclAllReduce + add.  kf_p2p_alloc creates this rank's symmetric buffer for messages of up to max_floats and returns its CUDA
 *      IPC handle (64 bytes); the caller gathers the handles of all ranks (torch.distributed / MPI) and passes them to kf_p2p_attach.
 *      kf_allreduce_residual: out = bf16(residual + bf16(sum over ranks of partial)), summed in rank order on every rank
 *      (bit-identical across ranks); falls back to kf_allreduce_f32 + kf_residual_add_f32 when the peer buffers are not attached
 *      or the message is larger than max_floats. ---- */
int kf_p2p_alloc(kf_ctx* ctx, size_t max_floats, int world, void* handle_out_64_bytes);
int kf_p2p_attach(kf_ctx* ctx, const void* handles_world_x_64_bytes, int rank, int world);
int kf_p2p_ready(kf_ctx* ctx);
int kf_p2p_release(kf_ctx* ctx); /* drop the peer buffers: kf_allreduce_residual then takes the NCCL path (all ranks must agree) */
int kf_allreduce_residual(kf_ctx* ctx, void* out_bf16_dev, const void* residual_bf16_dev, const float* partial_f32_dev, size_t n);

/* ---- the same exchange WITHOUT a launch of its own (kf_tp.cuh): for decode steps of up to 8 tokens it is the epilogue of the row-parallel
 *      matmul.  kf_linear_exchange = this rank's K-shard of O / down (SLP::Forw of proj_cat / down followed by the residual add,
 *      src/Device/CUDA/QKV.cu:676-688, NeuronFuse.cu:640-652): the epilogue pushes the fp32 partial rows into every peer's slot
 *      (flag-in-data stores over NVLink: no fence, no separate flag), adds the partials of all ranks in rank order, rounds, adds the
 *      residual and writes y -- y = bf16(residual + bf16(sum over ranks)), identical on every rank; y may alias residual.
 *      kf_tp_begin: first call of every forward on every rank (advances the device epoch counter, resets the exchange ordinal).
 *      kf_exchange_fused_ready: 1 when the path is available (peer buffers attached, knob tp_fused, M <= 8, M x cols within a slot). */
int kf_tp_begin(kf_ctx* ctx);
int kf_exchange_fused_ready(kf_ctx* ctx, int M, int cols);
int kf_linear_exchange(kf_ctx* ctx, void* y_dev, const kf_tensor_desc* w, const void* x_dev, int M, const void* residual_dev);

#ifdef __cplusplus
}
#endif
#endif /* KF_DEVICE_H */
