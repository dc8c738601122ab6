/*
 * koifish_oracle.h -- CPU ORACLE for the quantized-inference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a plain C++ restatement of the reference algorithm (gruai/koifish).  It may be
 * imported, linked or executed ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs -- always as the checker, never as the product.  Nothing under koifish_b200/
 * links it.
 *
 * Parity pinning: the reference ships NO golden vectors / known-answer tests for this path
 * (SURVEY.md section 8c).  The bit-layout half of the oracle (pack/unpack of 128-bit words) is pinned
 * against the reference's OWN macros compiled from /root/reference/src/PackedQ.hpp into
 * oracle/_ref/libkoifish_ref.so (see oracle/ref_shim.cpp, tests/test_oracle_vs_ref.py) and against the
 * fixtures generated from those macros under tests/golden/.  The arithmetic half (RTN_x / YinYang
 * quantiser, bf16 dequant, op order of the decode step) has no reference-run output to pin against
 * (the reference's quantiser only links together with its CUDA runtime): "parity unpinned" for those
 * functions beyond the invariants the reference asserts in-code (round trip, code range).
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

/* ---- dequant (src/Device/CUDA/T.cu:245-294 CU_Q128toX_) : w = RN_bf16(RN_bf16(step*(code-qbias)) - zero) ---- */
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

#ifdef __cplusplus
}
#endif
