/*
 * koifish_oracle.h -- CPU ORACLE for the quantized-inference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a plain C++ restatement of the reference algorithm (gruai/koifish).  It may be
 * imported, linked or executed ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs -- always as the checker, never as the product.  Nothing under koifish_b200/
 * links it.
 *
 * Parity pinning: the reference ships NO golden vectors / known-answer tests for this path, so the oracle
 * is pinned against the reference's own code compiled where it lies under /root/reference (recipe: oracle/Makefile):
 *   - 128-bit word layout: the PACK_/UNPACK_ macros of src/PackedQ.hpp (oracle/ref_shim.cpp -> oracle/_ref/libkoifish_ref.so,
 *     tests/test_oracle_vs_ref.py, fixtures tests/golden/packq_ref.npz);
 *   - fp32 matvec / rmsnorm: src/Utils/GST_float.cpp primitives (same library);
 *   - dequant arithmetic (CU_Q128toX_), the GPU quantise+pack kernels (CU_XtoQ128_, CU_XtoYYang_), RMSNorm (rms_norm_kernel,
 *     CU_rmsnorm_multihead), RoPE (CU_rope2_v0), decode attention (attention_qk_kernel / CU_softmax_multihead / attention_v_kernel)
 *     and the E5M2 byte codec: the reference's OWN CUDA kernels, compiled for sm_100a from src/Device/CUDA/T.cu and the headers it
 *     includes (oracle/ref_kernels.cu -> oracle/_ref/libkoifish_refgpu*.so) and run on the B200 next to koifish_b200's kernels
 *     (tests/test_gpu_refkernels.py, -m gpu);
 *   - NormalFloat4 and vendor-AWQ dequant: CU_Q42X_NF4 / CU_Q42X_awq from src/Device/CUDA/kernel/quantizer.cu (oracle/ref_kernels_q.cu ->
 *     oracle/_ref/libkoifish_refq.so; tests/test_gpu_kernels.py);
 *   - outputs of all of these reference kernels on seeded inputs are committed as tests/golden/refgpu_golden.npz (generator
 *     tests/golden/make_golden_refgpu.py, inputs tests/golden_cases.py) and checked against the oracle WITHOUT a GPU by
 *     tests/test_oracle_golden_refgpu.py.
 *   - the CPU packers GeQuant::RTN_x (4- / 2-bit asymmetric, symmetric, ternary), GeQuant::YinYang (1-bit) and RT_NormalF / _row_lut
 *     (NormalFloat4) with BIT_SET_k: the reference's own src/Tensor/GeQuant.cpp, src/Tensor/GTensor.cpp and src/Utils/CLI_params.cpp compiled
 *     where they lie and driven by oracle/ref_cpu_quant.cpp (-> oracle/_ref/libkoifish_refcpu.so; the rest of the framework those files
 *     mention is bound to 0 at link time and never reached).  kfo_quantize / kfo_nf4_quantize produce the SAME bytes and gama, bit for bit
 *     (tests/test_oracle.py: live against the library, and against tests/golden/refcpu_quant.npz generated from it by
 *     tests/golden/make_golden_refcpu.py);
 *   - the sampler port kfo_sample(selection = 1) against GeneratOnPrompt::Sample / LogitsInfo of src/Manifold/GoPT.cpp in the same library
 *     (tests/test_oracle.py);
 *   - the same library runs QUANT_CARD::Init4Neuron, QUANT_CARD::Vendor2JSONx and CHAT_SAMPLER::toChatML / InitPrefillTemplate of the reference:
 *     the product's quantizer-card selection, HF quantization_config mapping and ChatML templates are compared with them field by field /
 *     byte by byte (tests/test_cabi_host.py, tests/test_tokenizer.py);
 *   - the reference's tokenizer (src/TokenSet/HF_Tokenizer.cpp + vendored oniguruma / utf8proc) in oracle/_ref/libkoifish_reftok.so
 *     (oracle/ref_tokenizer.cpp) against csrc/TokenSet (tests/test_tokenizer.py);
 *   - the reference's checkpoint reader (src/Manifold/Serialize.cpp, src/Tensor/Safetensors.cpp) in oracle/_ref/libkoifish_refkun.so
 *     (oracle/ref_kun.cpp) reads the fish.kun files csrc/Tensor/KunFile.cpp writes (tests/test_kun_host.py);
 *   - the AWQ nibble order / values also against the reference's Python unpack (src/Python/test_awq.py, tests/golden/awq_ref_py.npz).
 * Still unpinned: the cuBLASLt GEMM (closed source; fp32 accumulation, order unspecified => tolerance).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* quantisation modes (QUANT_CARD, src/CLI_params.hpp:509-554 ; GeQuant ctor src/Tensor/GeQuant.cpp:107-124) */
enum kfo_qmode {
    KFO_RTN_ASYM = 0, /* {"quant_method":"RTN"}: qMin=0,qMax=2^b-1,qBias=0 */
    KFO_RTN_SYM  = 1, /* isSymmetric: qMin=-2^(b-1), qMax=2^(b-1)-1, qBias=-qMin */
    KFO_NF4      = 3, /* {"bits":4} without a quant_method: QUANT_MODE::RTNf, NormalFloat4 with a per-row codebook (kfo_nf4_*) */
    KFO_YYANG    = 2, /* {"quant_method":"yyang"}: bits 2 -> ternary {-1,0,1}+1 ; bits 1 -> {0,1} */
};

typedef struct {
    int qmin, qmax, qbias;
} kfo_qrange;

/* ---- numeric helpers ---- */
uint16_t kfo_f32_to_bf16(float f);  /* round-to-nearest-even, as __float2bfloat16_rn */
float kfo_bf16_to_f32(uint16_t h);
uint16_t kfo_bf16_mul(uint16_t a, uint16_t b); /* RN_bf16(a*b)  == device __hmul   */
uint16_t kfo_bf16_sub(uint16_t a, uint16_t b); /* RN_bf16(a-b)  == device __hsub (single rounding) */

/* ---- synthetic weights: counter-based, bit-reproducible on CPU and GPU (our generator; the reference uses
 * cuRAND on the GPU, src/Device/CUDA/huTensor.cu:199-210, which is not reproducible on a CPU) ---- */
void kfo_fill_normal(uint16_t* out_bf16, size_t n, uint64_t seed, float sigma, float mean);

/* ---- quantiser (src/Tensor/GeQuant.cpp:107-124, 375-404, 428-533 RTN_x, 536-628 YinYang) ---- */
int kfo_qrange_of(int bits, int mode, kfo_qrange* out);
size_t kfo_gama_elems(int rows, int cols, int group); /* rows + cols + 2*nGroup  (GeQuant.cpp:518) */
/* w: bf16 [rows, cols] row-major.  data_out: rows*cols*bits/8 bytes.  gama_out: kfo_gama_elems() bf16
 * laid out [R_SCALE rows][C_SCALE cols][ZERO nG][STEP nG] (src/Tensor/GTensor.cpp:456-510); R/C are left 0. */
int kfo_quantize(const uint16_t* w, int rows, int cols, int bits, int group, int mode, uint8_t* data_out, uint16_t* gama_out);

/* ---- 128-bit word pack/unpack (src/PackedQ.hpp:28-60, 99-239) ---- */
int kfo_pack_codes(const int32_t* codes, size_t n, int bits, uint8_t* data_out);
int kfo_unpack_codes(const uint8_t* data, size_t n, int bits, int32_t* codes_out);

/* ---- dequant (src/Device/CUDA/T.cu:245-294 CU_Q128toX_).  fused (default, what nvcc's -fmad=true / -use_fast_math build of the
 *      reference computes on sm_90+): w = RN_bf16(step*(code-qbias) - zero) ; two roundings (-fmad=false or pre-sm_90 builds):
 *      w = RN_bf16(RN_bf16(step*(code-qbias)) - zero).  kfo_set_dequant_fma selects (process-wide; 1 = fused). ---- */
void kfo_set_dequant_fma(int fused);
int kfo_get_dequant_fma(void);
uint16_t kfo_bf16_fms(uint16_t a, uint16_t b, uint16_t c); /* RN_bf16(a*b - c), one rounding == device fma.rn.bf16(a, b, -c) */
int kfo_dequant(const uint8_t* data, const uint16_t* gama, int rows, int cols, int bits, int group, int qbias, uint16_t* out_bf16);

/* ---- 8-bit E5M2-by-truncation (src/Device/CUDA/kernel/packedN.cuh:80-96, src/g_float.hpp:355-379) ---- */
void kfo_f8e5m2_encode(const uint16_t* w_bf16, size_t n, uint8_t* out);
void kfo_f8e5m2_decode(const uint8_t* in, size_t n, uint16_t* out_bf16);

/* ---- linear (src/Device/CUDA/kernel/gemm.cu:93-214): y[t][o] = RN_bf16(sum_k w[o][k]*x[t][k]), fp32 accumulate ---- */
void kfo_linear(uint16_t* y, const uint16_t* w, const uint16_t* x, int M, int N, int K);
/* fp32 result (no output rounding) for tolerance studies */
void kfo_linear_f32(float* y, const uint16_t* w, const uint16_t* x, int M, int N, int K);

/* ---- small ops ---- */
/* src/Device/CUDA/kernel/layernorm.cuh:801-859 rms_norm_kernel: (x*rsqrt(mean(x^2)+eps))*w, fp32, RN bf16 */
void kfo_rmsnorm(uint16_t* out, const uint16_t* x, const uint16_t* w, int rows, int dim, float eps);
/* src/Device/CUDA/kernel/operator.cuh:735-772 CU_rope2_v0: half-split pairs (j, j+hd/2); in place on [heads, hd] */
void kfo_rope(uint16_t* v, int n_heads, int head_dim, int pos, float theta);
/* src/Device/CUDA/Activation.cu:86-93: out = (g*u)/(1+exp(-g)) */
void kfo_swiglu(uint16_t* out, const uint16_t* gate, const uint16_t* up, size_t n);
/* src/Device/CUDA/kernel/packedN.cuh:867-875 CU_add3: out = RN_bf16(float(a)+float(b)) (reference rounds stochastically) */
void kfo_add(uint16_t* out, const uint16_t* a, const uint16_t* b, size_t n);
/* src/Device/CUDA/kernel/operator.cuh:573-668, 252-277: GQA decode attention for ONE query token at position pos.
 * q [n_head, hd] bf16, kcache/vcache [max_seq, n_kv*hd] bf16 (one layer), out [n_head, hd] bf16.
 * score_bf16 = 0: scores/probabilities kept in fp32 (pipe path, src/Device/CUDA/Generate.cu:273);
 * score_bf16 = 1: scores buffer is bf16 as in the neuron path (src/Manifold/TGraph.cpp:123-124, QKV.cu:670). */
void kfo_attention_decode(uint16_t* out, const uint16_t* q, const uint16_t* kcache, const uint16_t* vcache, int pos, int n_head, int n_kv,
                          int head_dim, int score_bf16);

/* ---- whole model (op order: SURVEY.md appendix A.6; src/Device/CUDA/QKV.cu:617-706, NeuronFuse.cu:615-656, 842-862,
 *      Generate.cu:180-346) ---- */
typedef struct {
    int n_layer, n_embd, n_ff, n_head, n_kv_head, head_dim, vocab, max_seq;
    float rope_theta, rms_eps;
    int tie_embed;
    int attn_bits, attn_mode; /* bits 16 = bf16 (no quant), 8 = F8E5M2, 4/2/1 = packed */
    int mlp_bits, mlp_mode;
    int embed_bits, embed_mode;
    int group;
    uint64_t seed;
    float sigma;      /* weight std (0.02 in the reference, huTensor.cu:204) */
    float norm_sigma; /* 0 => norm weights are exactly 1 (reference FIX_1); >0 => 1 + norm_sigma*z (tests) */
    int score_bf16;
} kfo_model_config;

typedef struct kfo_model kfo_model;
kfo_model* kfo_model_create(const kfo_model_config* cfg);
void kfo_model_destroy(kfo_model* m);
void kfo_model_reset(kfo_model* m);
/* run one token at position pos (appends K/V at pos); logits_out: vocab bf16 (may be NULL for prefill tokens) */
int kfo_model_forward(kfo_model* m, int token, int pos, uint16_t* logits_out);
/* one transformer layer only (bench sample): x in/out [n_embd] bf16 */
int kfo_model_layer(kfo_model* m, int layer, int pos, uint16_t* x_inout);
/* tensor ids: see kfo_tensor_seed(); returns the dequantised bf16 weight the oracle uses (debug / parity) */
const uint16_t* kfo_model_weight(kfo_model* m, int tensor_id, size_t* n_out);
const uint16_t* kfo_model_kcache(kfo_model* m, int layer);
const uint16_t* kfo_model_vcache(kfo_model* m, int layer);
uint64_t kfo_tensor_seed(uint64_t model_seed, int tensor_id);
int kfo_num_threads(void);
/* NormalFloat4 (QUANT_MODE::RTNf): GeQuant::RT_NormalF / _row_lut src/Tensor/GeQuant.cpp:696-748, Distri_PIPE::Prepare :650-682, X2NormalF
 * :684-700, NF4_LUT src/g_float.hpp:542-558, BIT_SET_k src/Utils/CLI_params.cpp:2177-2191; dequant CU_Q42X_NF4
 * src/Device/CUDA/kernel/quantizer.cu:612-654.  data: rows*cols/2 bytes; gama: rows + cols + 16*rows bf16. */
/* Vendor AWQ layout as the reference reads it: CU_Q42X_awq src/Device/CUDA/kernel/quantizer.cu:132-156, CU_I2Q4_unpack + AWQ_REVERSE_ORDER
 * src/Device/CUDA/kernel/packedN.cuh:109-116.  Stored matrix [M][N] (M = in_features, N = out_features: SLP::Forw uses transA = 0 for AWQ,
 * NeuronFuse.cu:305-381): qweight int32 [M][N/8], qzeros int32 [M/128][N/8], scales fp16 [M/128][N]; element 8c+k of a row sits in nibble
 * AWQ_REVERSE_ORDER[k] of word c;  w = bf16( float(q - z) * float(scale) ).  out: bf16 [M][N]. */
int kfo_awq_dequant(const uint32_t* qweight, const uint32_t* qzeros, const uint16_t* scales_f16, int M, int N, uint16_t* out);
/* test helper (the reference has no AWQ quantiser: AWQ tensors come from vendor checkpoints): RTN-asymmetric codes per (128 rows, column) */
int kfo_awq_pack(const uint16_t* w_bf16_MN, int M, int N, uint32_t* qweight, uint32_t* qzeros, uint16_t* scales_f16);
int kfo_nf4_quantize(const uint16_t* w, int rows, int cols, uint8_t* data_out, uint16_t* gama_out);
int kfo_nf4_dequant(const uint8_t* data, const uint16_t* gama, int rows, int cols, uint16_t* out);
/* GeneratOnPrompt::Sample (src/Manifold/GoPT.cpp:614-630): returns the token, advances *rng_state; selection 0 = true top-k, 1 = the
 * reference's TOPK_heap::Select as written */
int kfo_sample(const uint16_t* logits, int vocab, float temperature, int top_k, float top_p, uint64_t* rng_state, int selection, int* n_pick_out);

#ifdef __cplusplus
}
#endif
