/*
 * koifish_oracle.cpp -- CPU ORACLE (test infrastructure only; see koifish_oracle.h for the usage rules and the
 * "parity unpinned" statement).  Plain C++17 + OpenMP restatement of the reference algorithm; every function
 * cites the reference file:line it follows (paths relative to /root/reference).  No reference source is copied.
 */
#include "koifish_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cfloat>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__F16C__)
#include <immintrin.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * bf16 arithmetic.  floatGama == floatX == __nv_bfloat16 in the reference (src/g_float.hpp:246-261), and the
 * dequant kernel evaluates (step * (floatGama)k - zero) with bf16 operators, i.e. one rounding per operator
 * (src/Device/CUDA/T.cu:274).
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
extern "C" float kfo_bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }
extern "C" uint16_t kfo_f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    if ((u & 0x7fffffffu) > 0x7f800000u)
        return (uint16_t)((u >> 16) | 0x0040u); /* quiet NaN */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb; /* round to nearest even */
    return (uint16_t)(u >> 16);
}
/* double -> bf16 with a single rounding: go through float with round-to-odd (float keeps 16 more bits than bf16,
 * so RN(round_to_odd(x)) == RN(x)). */
static inline uint16_t f64_to_bf16(double d) {
    float f = (float)d;
    if ((double)f != d && !isnan(d) && !isinf(f)) {
        if (fabs((double)f) > fabs(d))
            f = nextafterf(f, 0.0f);   /* truncate toward zero */
        f = u2f(f2u(f) | 1u);          /* sticky bit */
    }
    return kfo_f32_to_bf16(f);
}
extern "C" uint16_t kfo_bf16_mul(uint16_t a, uint16_t b) {
    /* 8-bit x 8-bit significands: the product is exact in fp32 unless it under/overflows fp32's range, so use double */
    return f64_to_bf16((double)kfo_bf16_to_f32(a) * (double)kfo_bf16_to_f32(b));
}
extern "C" uint16_t kfo_bf16_sub(uint16_t a, uint16_t b) { return f64_to_bf16((double)kfo_bf16_to_f32(a) - (double)kfo_bf16_to_f32(b)); }

/* ------------------------------------------------------------------------------------------------
 * Synthetic weights.  Reference: N(0, 0.02^2) from cuRAND on the device (src/Device/CUDA/huTensor.cu:199-210).  cuRAND
 * streams cannot be reproduced on a CPU, so the framework defines its own counter-based generator (integer hash +
 * Irwin-Hall sum of four 16-bit uniforms, one fp32 fma) that is bit-identical on CPU and GPU.  The product's device
 * kernel (koifish_b200/csrc/Device/fill.cu) implements the same definition independently.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float synth_z(uint64_t seed, uint64_t idx) {
    uint64_t h = mix64(seed * 0xD1342543DE82EF95ull + idx);
    int s      = (int)(h & 0xffff) + (int)((h >> 16) & 0xffff) + (int)((h >> 32) & 0xffff) + (int)((h >> 48) & 0xffff);
    return (float)(s - 131070);
}
#define KFO_IH_STD 37837.227f /* std of the sum of four U{0..65535} */
extern "C" void kfo_fill_normal(uint16_t* out, size_t n, uint64_t seed, float sigma, float mean) {
    const float scale = sigma / KFO_IH_STD;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) out[i] = kfo_f32_to_bf16(fmaf(synth_z(seed, i), scale, mean));
}
extern "C" uint64_t kfo_tensor_seed(uint64_t model_seed, int tensor_id) { return model_seed * 1000003ull + (uint64_t)tensor_id; }

/* ------------------------------------------------------------------------------------------------
 * Code ranges: GeQuant ctor, src/Tensor/GeQuant.cpp:107-124
 * ---------------------------------------------------------------------------------------------- */
extern "C" int kfo_qrange_of(int bits, int mode, kfo_qrange* r) {
    if (mode == KFO_YYANG) {
        if (bits == 2) {
            r->qmax = 1, r->qmin = -1, r->qbias = 1;
        } else if (bits == 1) {
            r->qmax = 1, r->qmin = 0, r->qbias = 0;
        } else
            return -1;
    } else if (mode == KFO_RTN_SYM) {
        r->qmin  = -(1 << (bits - 1));
        r->qmax  = (1 << (bits - 1)) - 1;
        r->qbias = -r->qmin;
    } else {
        r->qmin = 0, r->qmax = (1 << bits) - 1, r->qbias = 0;
    }
    return 0;
}
extern "C" size_t kfo_gama_elems(int rows, int cols, int group) { return (size_t)rows + cols + 2 * ((size_t)rows * cols / group); }

/* ------------------------------------------------------------------------------------------------
 * 128-bit words: src/PackedQ.hpp:28-31 (struct order {low, high} => bytes 0-7 = low, 8-15 = high),
 * :99-141 PACK_4to128_, :185-198 PACK_2to128_, :200-211 PACK_1to128_, and the UNPACK_* mirrors :143-239.
 * Code j of a word sits in `high` when j < half, at shift 64 - bits*(j+1); else in `low` at 64 - bits*(j-half+1).
 * ---------------------------------------------------------------------------------------------- */
static inline void pack_word(const int32_t* codes, int bits, uint8_t* dst16) {
    const int per = 128 / bits, half = per / 2;
    const uint64_t mask = (1ull << bits) - 1;
    uint64_t high = 0, low = 0;
    for (int j = 0; j < half; j++) {
        high |= ((uint64_t)codes[j] & mask) << (64 - bits * (j + 1));
        low |= ((uint64_t)codes[j + half] & mask) << (64 - bits * (j + 1));
    }
    memcpy(dst16, &low, 8); /* little-endian host, as the reference assumes (GTensor.cpp:512-514) */
    memcpy(dst16 + 8, &high, 8);
}
static inline void unpack_word(const uint8_t* src16, int bits, int32_t* codes) {
    const int per = 128 / bits, half = per / 2;
    const uint64_t mask = (1ull << bits) - 1;
    uint64_t high, low;
    memcpy(&low, src16, 8);
    memcpy(&high, src16 + 8, 8);
    for (int j = 0; j < half; j++) {
        codes[j]        = (int32_t)((high >> (64 - bits * (j + 1))) & mask);
        codes[j + half] = (int32_t)((low >> (64 - bits * (j + 1))) & mask);
    }
}
extern "C" int kfo_pack_codes(const int32_t* codes, size_t n, int bits, uint8_t* out) {
    if (bits != 4 && bits != 2 && bits != 1)
        return -1;
    const size_t per = 128 / bits;
    if (n % per)
        return -2;
    for (size_t w = 0; w < n / per; w++) pack_word(codes + w * per, bits, out + 16 * w);
    return 0;
}
extern "C" int kfo_unpack_codes(const uint8_t* data, size_t n, int bits, int32_t* out) {
    if (bits != 4 && bits != 2 && bits != 1)
        return -1;
    const size_t per = 128 / bits;
    if (n % per)
        return -2;
    for (size_t w = 0; w < n / per; w++) unpack_word(data + 16 * w, bits, out + w * per);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Quantiser.  RTN_x: src/Tensor/GeQuant.cpp:428-533 ; YinYang (1-bit): :536-628 ; groups: :375-404
 * (runs of T_group consecutive elements of the row-major matrix) ; gama layout: src/Tensor/GTensor.cpp:456-510.
 * All arithmetic in float exactly as written there (vSum in double); ZERO/STEP are stored as bf16 (RN) while the
 * codes come from the float zero/step.
 * Deviation: a constant group (vmax==vmin => step 0) divides by zero in the reference (UB); the oracle emits code
 * qbias (k = 0) for such a group.
 * ---------------------------------------------------------------------------------------------- */
extern "C" int kfo_quantize(const uint16_t* w, int rows, int cols, int bits, int group, int mode, uint8_t* data_out, uint16_t* gama_out) {
    kfo_qrange qr;
    if (kfo_qrange_of(bits, mode, &qr))
        return -1;
    if (bits != 4 && bits != 2 && bits != 1)
        return -1;
    if (bits == 1 && mode != KFO_YYANG)
        return -1; /* Core(): bits==1 -> YinYang only (GeQuant.cpp:909) */
    const int per128 = 128 / bits;
    const size_t nElem = (size_t)rows * cols;
    if (nElem % group || group % per128)
        return -2; /* assert(nCol % nPer128 == 0), GeQuant.cpp:438 */
    const size_t nG = nElem / group;
    memset(gama_out, 0, sizeof(uint16_t) * ((size_t)rows + cols)); /* R/C scales unused (NO_NORMAL, GeQuant.cpp:844-852) */
    uint16_t* gZero = gama_out + rows + cols;
    uint16_t* gStep = gZero + nG;
    const int qMin = qr.qmin, qMax = qr.qmax, qBias = qr.qbias;
#pragma omp parallel for schedule(static)
    for (long long g = 0; g < (long long)nG; g++) {
        std::vector<float> tmp(group);
        float vmax = -FLT_MAX, vmin = FLT_MAX;
        double vSum = 0.0;
        const uint16_t* dat = w + (size_t)g * group;
        for (int i = 0; i < group; i++) {
            float a = kfo_bf16_to_f32(dat[i]); /* sR = sC = 1 */
            tmp[i]  = a;
            vmax = std::max(vmax, a), vmin = std::min(vmin, a);
            if (bits == 1)
                vSum += a < 0.0 ? 0.0 : a * a; /* YinYang energy, GeQuant.cpp:573 (a*a in float, summed in double) */
            else
                vSum += fabsf(a);              /* RTN_x, GeQuant.cpp:461 */
        }
        float step, zero;
        if (bits == 1) {
            float vMean = (float)sqrt(vSum / group); /* GeQuant.cpp:576 */
            step = std::max(1e-5f, vMean), zero = 0;
        } else {
            float vMean = (float)(vSum / group);
            step = (vmax - vmin) / (float)(qMax - qMin), zero = -vmin; /* GeQuant.cpp:465 */
            if (mode == KFO_YYANG) {
                step = std::max(1e-5f, vMean), zero = 0;               /* :466-468 */
            } else if (mode == KFO_RTN_SYM) {
                step = std::max(fabsf(vmax), fabsf(vmin)) / (float)qMax, zero = 0; /* :470-471 */
            }
        }
        gZero[g] = kfo_f32_to_bf16(zero), gStep[g] = kfo_f32_to_bf16(step); /* :477 implicit float -> bf16 */
        int32_t qq[128];
        uint8_t* quanti = data_out + (size_t)g * group * bits / 8;
        for (int i = 0; i < group / per128; i++) {
            for (int pos = 0; pos < per128; pos++) {
                float a = tmp[pos + i * per128];
                int qid;
                if (step == 0.0f)
                    qid = 0;
                else
                    qid = (int)roundf((a + zero) / step); /* std::round: half away from zero, :484 */
                if (mode == KFO_YYANG || bits == 1) {
                    qid = std::max(qid, qMin), qid = std::min(qid, qMax); /* clamped only for yyang, :485-486 / :586-587 */
                }
                qq[pos] = qid + qBias;
            }
            pack_word(qq, bits, quanti + 16 * i);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Dequant: CU_Q128toX_, src/Device/CUDA/T.cu:245-294.  g0 = (step * (floatGama)(q - qBias) - zero) * sR with bf16
 * operators and sR = 1.  How many roundings that expression has depends on how the reference is BUILT, and it is pinned against
 * the reference's own kernel compiled both ways (oracle/ref_kernels.cu, tests/test_gpu_refkernels.py):
 *   fused (default)  nvcc's default -fmad=true (implied by the reference's -use_fast_math, CMakeLists.txt:141) contracts the bf16
 *                    multiply and subtract into ONE fma.rn.bf16 on sm_90+ (SASS: HFMA2.BF16 step, k, -zero):  w = RN_bf16(step*k - zero);
 *   two roundings    -fmad=false, and every pre-sm_90 build (there the bf16 operators go through fp32 with a rounding after each
 *                    operator):  p = RN_bf16(step*k) ; w = RN_bf16(p - zero).
 * ---------------------------------------------------------------------------------------------- */
static int g_dequant_fma = 1;
extern "C" void kfo_set_dequant_fma(int fused) { g_dequant_fma = fused ? 1 : 0; }
extern "C" int kfo_get_dequant_fma(void) { return g_dequant_fma; }
/* RN_bf16(a*b - c) with a single rounding: the product of two bf16 numbers is exact in double, and so is the difference unless the
 * exponents are > 2^29 apart (never for quantiser scales) */
extern "C" uint16_t kfo_bf16_fms(uint16_t a, uint16_t b, uint16_t c) {
    return f64_to_bf16((double)kfo_bf16_to_f32(a) * (double)kfo_bf16_to_f32(b) - (double)kfo_bf16_to_f32(c));
}
extern "C" int kfo_dequant(const uint8_t* data, const uint16_t* gama, int rows, int cols, int bits, int group, int qbias, uint16_t* out) {
    if (bits != 4 && bits != 2 && bits != 1)
        return -1;
    const int per128 = 128 / bits;
    const size_t nElem = (size_t)rows * cols;
    if (nElem % group || group % per128)
        return -2;
    const size_t nG = nElem / group;
    const uint16_t* gZero = gama + rows + cols;
    const uint16_t* gStep = gZero + nG;
#pragma omp parallel for schedule(static)
    for (long long g = 0; g < (long long)nG; g++) {
        int32_t qq[128];
        const uint16_t zero = gZero[g], step = gStep[g];
        for (int i = 0; i < group / per128; i++) {
            unpack_word(data + ((size_t)g * group * bits / 8) + 16 * i, bits, qq);
            for (int pos = 0; pos < per128; pos++) {
                uint16_t kq = kfo_f32_to_bf16((float)(qq[pos] - qbias)); /* small integers are exact in bf16 */
                out[(size_t)g * group + i * per128 + pos] =
                    g_dequant_fma ? kfo_bf16_fms(step, kq, zero) : kfo_bf16_sub(kfo_bf16_mul(step, kq), zero);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * 8-bit: CU_16BF2T<f8e5> src/Device/CUDA/kernel/packedN.cuh:80-96 (bf16 -> float -> fp16 RN -> keep the HIGH byte),
 * T2Float<f8e5> src/g_float.hpp:355-379 (byte<<8 read as fp16), CU_F82Float operator.cuh:535-543 (-> bf16).
 * ---------------------------------------------------------------------------------------------- */
static inline uint16_t f32_to_f16_rn(float f) {
#if defined(__F16C__)
    return (uint16_t)_cvtss_sh(f, 0);
#else
    /* portable RN conversion */
    uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u;
    int32_t e  = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff)
        return (uint16_t)(sign | 0x7c00u | (m ? 0x200u : 0));
    if (e >= 31)
        return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10)
            return (uint16_t)sign;
        m |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e), half = 1u << (shift - 1), r = m >> shift, rem = m & ((1u << shift) - 1);
        if (rem > half || (rem == half && (r & 1)))
            r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1)))
        r++;
    return (uint16_t)(sign | r);
#endif
}
static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu;
    if (e == 0) {
        if (m == 0)
            return u2f(sign);
        float v = (float)m * 5.9604644775390625e-8f; /* m * 2^-24 */
        return (h & 0x8000u) ? -v : v;
    }
    if (e == 31)
        return u2f(sign | 0x7f800000u | (m << 13));
    return u2f(sign | ((e - 15 + 127) << 23) | (m << 13));
}
extern "C" void kfo_f8e5m2_encode(const uint16_t* w, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)(f32_to_f16_rn(kfo_bf16_to_f32(w[i])) >> 8);
}
extern "C" void kfo_f8e5m2_decode(const uint8_t* in, size_t n, uint16_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = kfo_f32_to_bf16(f16_to_f32((uint16_t)((uint16_t)in[i] << 8)));
}

/* ------------------------------------------------------------------------------------------------
 * Linear: CU_mm_blasLt, src/Device/CUDA/kernel/gemm.cu:93-214 (CUBLAS_COMPUTE_32F, bf16 in/out, :124-126, 196-202);
 * TASKA_AxB src/Tensor/GTensor.hpp:703-741.  d is [tokens][OC] row-major, b is [tokens][IC].  Accumulation order is
 * unspecified in the reference (cuBLASLt heuristics) -> parity by tolerance on y.
 * ---------------------------------------------------------------------------------------------- */
static inline float dot_bf16_f32(const uint16_t* w, const float* x, int K) {
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (int k = 0; k < K; k++) acc += u2f((uint32_t)w[k] << 16) * x[k];
    return acc;
}
extern "C" void kfo_linear_f32(float* y, const uint16_t* w, const uint16_t* x, int M, int N, int K) {
    std::vector<float> xf((size_t)M * K);
    for (size_t i = 0; i < xf.size(); i++) xf[i] = kfo_bf16_to_f32(x[i]);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        for (int m = 0; m < M; m++) y[(size_t)m * N + n] = dot_bf16_f32(w + (size_t)n * K, xf.data() + (size_t)m * K, K);
    }
}
extern "C" void kfo_linear(uint16_t* y, const uint16_t* w, const uint16_t* x, int M, int N, int K) {
    std::vector<float> yf((size_t)M * N);
    kfo_linear_f32(yf.data(), w, x, M, N, K);
    for (size_t i = 0; i < yf.size(); i++) y[i] = kfo_f32_to_bf16(yf[i]);
}

/* ------------------------------------------------------------------------------------------------
 * Small ops
 * ---------------------------------------------------------------------------------------------- */
/* rms_norm_kernel, src/Device/CUDA/kernel/layernorm.cuh:801-859 (and CU_rmsnorm_multihead :750-798 for per-head rows):
 * fp32 sum of squares, rsqrtf(sum/D + eps), (x*s)*w, RN to bf16. */
extern "C" void kfo_rmsnorm(uint16_t* out, const uint16_t* x, const uint16_t* w, int rows, int dim, float eps) {
    for (int r = 0; r < rows; r++) {
        const uint16_t* xr = x + (size_t)r * dim;
        float ss = 0.f;
        for (int i = 0; i < dim; i++) {
            float v = kfo_bf16_to_f32(xr[i]);
            ss      = fmaf(v, v, ss);
        }
        float s = 1.0f / sqrtf(fmaf(ss, 1.0f / (float)dim, eps));
        for (int i = 0; i < dim; i++) out[(size_t)r * dim + i] = kfo_f32_to_bf16((kfo_bf16_to_f32(xr[i]) * s) * kfo_bf16_to_f32(w[i]));
    }
}
/* CU_rope2_v0, src/Device/CUDA/kernel/operator.cuh:735-772: pairs (j, j+hd/2), inv_freq = 1/powf(theta, 2j/hd),
 * angle = pos*inv_freq; out1 = r*cos - i*sin ; out2 = r*sin + i*cos ; bf16 RN (reference: stochastic rounding). */
extern "C" void kfo_rope(uint16_t* v, int n_heads, int head_dim, int pos, float theta) {
    const int half = head_dim / 2;
    for (int h = 0; h < n_heads; h++) {
        uint16_t* p = v + (size_t)h * head_dim;
        for (int j = 0; j < half; j++) {
            float inv_freq = 1.0f / powf(theta, (float)(j * 2) / (float)head_dim);
            float angle = (float)pos * inv_freq, c = cosf(angle), s = sinf(angle);
            float re = kfo_bf16_to_f32(p[j]), im = kfo_bf16_to_f32(p[j + half]);
            p[j]        = kfo_f32_to_bf16(fmaf(re, c, -(im * s)));
            p[j + half] = kfo_f32_to_bf16(fmaf(re, s, im * c));
        }
    }
}
/* CU_swiglu_v0, src/Device/CUDA/Activation.cu:86-93 */
extern "C" void kfo_swiglu(uint16_t* out, const uint16_t* gate, const uint16_t* up, size_t n) {
    for (size_t i = 0; i < n; i++) {
        float g = kfo_bf16_to_f32(gate[i]), u = kfo_bf16_to_f32(up[i]);
        out[i]  = kfo_f32_to_bf16((g * u) / (1.0f + expf(-g)));
    }
}
/* CU_add3 / PackedN::Add2, src/Device/CUDA/kernel/packedN.cuh:867-875, 446-453 */
extern "C" void kfo_add(uint16_t* out, const uint16_t* a, const uint16_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = kfo_f32_to_bf16(kfo_bf16_to_f32(a[i]) + kfo_bf16_to_f32(b[i]));
}
/* attention_qk_kernel + CU_softmax_multihead + attention_v_kernel, src/Device/CUDA/kernel/operator.cuh:573-632,
 * 252-277, 650-668.  kv head = h / (n_head/n_kv); scale = 1/sqrt(hd) applied as a division. */
extern "C" void kfo_attention_decode(uint16_t* out, const uint16_t* q, const uint16_t* kc, const uint16_t* vc, int pos, int n_head, int n_kv,
                                     int hd, int score_bf16) {
    const int kv_mul = n_head / n_kv, kv_dim = n_kv * hd, len = pos + 1;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_head; h++) {
        std::vector<float> att(len);
        const uint16_t* qh = q + (size_t)h * hd;
        const int kvh      = h / kv_mul;
        for (int t = 0; t < len; t++) {
            const uint16_t* k = kc + (size_t)t * kv_dim + (size_t)kvh * hd;
            float score = 0.f;
            for (int i = 0; i < hd; i++) score = fmaf(kfo_bf16_to_f32(qh[i]), kfo_bf16_to_f32(k[i]), score);
            score /= sqrtf((float)hd);
            att[t] = score_bf16 ? kfo_bf16_to_f32(kfo_f32_to_bf16(score)) : score;
        }
        float mx = -1e9f;
        for (int t = 0; t < len; t++) mx = std::max(mx, att[t]);
        float sum = 0.f;
        for (int t = 0; t < len; t++) {
            float d = att[t] - mx;
            if (score_bf16)
                d = kfo_bf16_to_f32(kfo_f32_to_bf16(d)); /* bf16 - bf16 operator, operator.cuh:268 with T=bf16 */
            float a = expf(d);
            sum += a;
            att[t] = score_bf16 ? kfo_bf16_to_f32(kfo_f32_to_bf16(a)) : a;
        }
        float inv = 1.0f / sum;
        for (int t = 0; t < len; t++) {
            if (score_bf16)
                att[t] = kfo_bf16_to_f32(kfo_bf16_mul(kfo_f32_to_bf16(att[t]), kfo_f32_to_bf16(inv))); /* bf16 *= float, :275 */
            else
                att[t] *= inv;
        }
        for (int i = 0; i < hd; i++) {
            float acc = 0.f;
            for (int t = 0; t < len; t++) acc = fmaf(att[t], kfo_bf16_to_f32(vc[(size_t)t * kv_dim + (size_t)kvh * hd + i]), acc);
            out[(size_t)h * hd + i] = kfo_f32_to_bf16(acc);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Whole model.  Tensor ids (also the seed derivation the product uses for synthetic weights):
 *   0 embed_tokens [vocab,E] ; 1 final norm [E] ; 2 lm_head [vocab,E] (absent when tied)
 *   layer l: base = 16 + 16*l : +0 input_layernorm, +1 q_proj [H*hd,E], +2 k_proj [KV*hd,E], +3 v_proj, +4 q_norm [hd],
 *            +5 k_norm [hd], +6 o_proj [E,H*hd], +7 post_attention_layernorm, +8 gate [F,E], +9 up [F,E], +10 down [E,F]
 * Weight init: huTensor::InitParam, src/Device/CUDA/huTensor.cu:157-231 (N(0,0.02^2), norms FIX_1); quantise-at-load:
 * GeQuant::LowBit_worker src/Tensor/GeQuant.cpp:830-905.
 * ---------------------------------------------------------------------------------------------- */
struct kfo_model {
    kfo_model_config c;
    std::vector<std::vector<uint16_t>> w; /* dequantised bf16 weights by tensor id */
    std::vector<uint16_t> kc, vc;         /* [L][max_seq][kv_dim] : KVCache, src/Utils/Cache.cpp:14-60 */
};
/* ------------------------------------------------------------------------------------------------
 * Vendor AWQ layout: CU_Q42X_awq (quantizer.cu:132-156) with CU_I2Q4_unpack (packedN.cuh:109-116).
 * ---------------------------------------------------------------------------------------------- */
static const int kfo_awq_order[8] = {0, 4, 1, 5, 2, 6, 3, 7}; /* AWQ_REVERSE_ORDER: element k of a word sits at nibble order[k] */
static inline float kfo_f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            uint32_t m = man;
            do { e++, m <<= 1; } while (!(m & 0x400u));
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (m & 0x3ffu) << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | man << 13;
    } else {
        bits = sign | (exp + 112) << 23 | man << 13;
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
extern "C" int kfo_awq_dequant(const uint32_t* qweight, const uint32_t* qzeros, const uint16_t* scales_f16, int M, int N, uint16_t* out) {
    if (N % 8 || M % 128) return -1;
#pragma omp parallel for schedule(static)
    for (int row = 0; row < M; row++) {
        const uint32_t* q = qweight + (size_t)row * (N / 8);
        const uint32_t* z = qzeros + (size_t)(row / 128) * (N / 8);
        const uint16_t* sc = scales_f16 + (size_t)(row / 128) * N;
        for (int c = 0; c < N / 8; c++)
            for (int k = 0; k < 8; k++) {
                const int sft = kfo_awq_order[k] * 4;
                const int qv = (q[c] >> sft) & 0xF, zv = (z[c] >> sft) & 0xF;
                const float g0 = (float)(qv - zv) * kfo_f16_to_f32(sc[8 * c + k]); /* (q8[i] - z8[i]) * (float)scale8[i] */
                out[(size_t)row * N + 8 * c + k] = kfo_f32_to_bf16(g0);
            }
    }
    return 0;
}
extern "C" int kfo_awq_pack(const uint16_t* w, int M, int N, uint32_t* qweight, uint32_t* qzeros, uint16_t* scales_f16) {
    if (N % 8 || M % 128) return -1;
    memset(qweight, 0, sizeof(uint32_t) * (size_t)M * (N / 8));
    memset(qzeros, 0, sizeof(uint32_t) * (size_t)(M / 128) * (N / 8));
    for (int g = 0; g < M / 128; g++)
        for (int col = 0; col < N; col++) {
            float vmin = FLT_MAX, vmax = -FLT_MAX;
            for (int r = 0; r < 128; r++) {
                const float a = kfo_bf16_to_f32(w[(size_t)(g * 128 + r) * N + col]);
                vmin = std::min(vmin, a), vmax = std::max(vmax, a);
            }
            float scale = (vmax - vmin) / 15.0f;
            if (!(scale > 0)) scale = 1.0f;
            /* fp16 scale, round to nearest via the float -> half conversion of the compiler */
            const _Float16 hs = (_Float16)scale;
            uint16_t hbits;
            memcpy(&hbits, &hs, 2);
            scales_f16[(size_t)g * N + col] = hbits;
            const float s = (float)hs;
            int zq = (int)lrintf(-vmin / s);
            zq     = std::max(0, std::min(15, zq));
            const int word = col / 8, sft = kfo_awq_order[col % 8] * 4;
            qzeros[(size_t)g * (N / 8) + word] |= (uint32_t)zq << sft;
            for (int r = 0; r < 128; r++) {
                const float a = kfo_bf16_to_f32(w[(size_t)(g * 128 + r) * N + col]);
                int qv        = (int)lrintf(a / s) + zq;
                qv            = std::max(0, std::min(15, qv));
                qweight[(size_t)(g * 128 + r) * (N / 8) + word] |= (uint32_t)qv << sft;
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * NormalFloat4, QUANT_MODE::RTNf -- what {"bits": 4} without a quant_method selects (GeQuant.cpp:1270-1280).
 * Quantise: GeQuant::_row_lut (GeQuant.cpp:696-732): per row, Distri_PIPE::Next over the fp32 values (vmin / vmax / abs_max, GTensor.hpp:141-148),
 * Prepare(16) in its default symmetric case (:674-681: scale = abs_max > 0 ? 1 / abs_max : 1 ; codebook[i] = table[i] / scale), the LUT stored
 * as bf16 (Float2T<floatGama>), the code = first minimum of |w - codebook[i]| over the FP32 codebook (X2NormalF :684-700), packed by
 * BIT_SET_k (CLI_params.cpp:2177-2191: MSB-first bit stream, element i at bit 4 i).  No row / column normalisation (NORMAL_MODE::NO_NORMAL).
 * Dequant: CU_Q42X_NF4 (quantizer.cu:612-654) with rc_normal = 0: w = lut[row][code] (id0 = high nibble = the even element).
 * ---------------------------------------------------------------------------------------------- */
static const float kfo_nf4_table[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f, -0.28444138169288635f,
                                        -0.18477343022823334f, -0.09105003625154495f, 0.0f, 0.07958029955625534f, 0.16093020141124725f,
                                        0.24611230194568634f, 0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f,
                                        0.7229568362236023f, 1.0f}; /* NF4_LUT::table, src/g_float.hpp:543-558 */
extern "C" int kfo_nf4_quantize(const uint16_t* w, int rows, int cols, uint8_t* data_out, uint16_t* gama_out) {
    if (cols % 2) return -1;
    memset(gama_out, 0, sizeof(uint16_t) * ((size_t)rows + cols));
    memset(data_out, 0, (size_t)rows * cols / 2);
#pragma omp parallel for schedule(static)
    for (int row = 0; row < rows; row++) {
        const uint16_t* wr = w + (size_t)row * cols;
        float vmin = FLT_MAX, vmax = -FLT_MAX;
        for (int i = 0; i < cols; i++) {
            const float a = kfo_bf16_to_f32(wr[i]);
            vmax = std::max(vmax, a), vmin = std::min(vmin, a);
        }
        const float abs_max = std::max(std::fabs(vmin), std::fabs(vmax));
        const float scale   = abs_max > 0 ? 1.0f / abs_max : 1.0f;
        float cb[16];
        uint16_t* lut = gama_out + rows + cols + (size_t)row * 16;
        for (int i = 0; i < 16; i++) cb[i] = kfo_nf4_table[i] / scale, lut[i] = kfo_f32_to_bf16(cb[i]);
        uint8_t* q = data_out + (size_t)row * cols / 2;
        for (int i = 0; i < cols; i++) {
            const float a = kfo_bf16_to_f32(wr[i]);
            float best    = std::numeric_limits<float>::max();
            int bi        = 0;
            for (int c = 0; c < 16; c++) {
                const float d = std::abs(a - cb[c]);
                if (d < best) best = d, bi = c;
            }
            size_t boff = (size_t)i * 4; /* BIT_SET_k */
            for (int b = 0; b < 4; b++, boff++)
                if ((bi >> (3 - b)) & 1) q[boff / 8] |= (uint8_t)(1u << (7 - boff % 8));
        }
    }
    return 0;
}
extern "C" int kfo_nf4_dequant(const uint8_t* data, const uint16_t* gama, int rows, int cols, uint16_t* out) {
    if (cols % 2) return -1;
#pragma omp parallel for schedule(static)
    for (int row = 0; row < rows; row++) {
        const uint16_t* lut = gama + rows + cols + (size_t)row * 16;
        const uint8_t* q    = data + (size_t)row * cols / 2;
        for (int k = 0; k < cols / 2; k++) {
            out[(size_t)row * cols + 2 * k]     = lut[(q[k] >> 4) & 0x0F];
            out[(size_t)row * cols + 2 * k + 1] = lut[q[k] & 0x0F];
        }
    }
    return 0;
}

static void make_weight(std::vector<uint16_t>& dst, int rows, int cols, int bits, int mode, int group, uint64_t seed, float sigma) {
    const size_t n = (size_t)rows * cols;
    dst.resize(n);
    kfo_fill_normal(dst.data(), n, seed, sigma, 0.f);
    if (bits == 16)
        return;
    if (bits == 8) {
        std::vector<uint8_t> q(n);
        kfo_f8e5m2_encode(dst.data(), n, q.data());
        kfo_f8e5m2_decode(q.data(), n, dst.data());
        return;
    }
    if (mode == KFO_NF4) {
        std::vector<uint8_t> data(n / 2);
        std::vector<uint16_t> gama((size_t)rows + cols + 16 * (size_t)rows);
        kfo_nf4_quantize(dst.data(), rows, cols, data.data(), gama.data());
        kfo_nf4_dequant(data.data(), gama.data(), rows, cols, dst.data());
        return;
    }
    kfo_qrange qr;
    kfo_qrange_of(bits, mode, &qr);
    std::vector<uint8_t> data(n * bits / 8);
    std::vector<uint16_t> gama(kfo_gama_elems(rows, cols, group));
    kfo_quantize(dst.data(), rows, cols, bits, group, mode, data.data(), gama.data());
    kfo_dequant(data.data(), gama.data(), rows, cols, bits, group, qr.qbias, dst.data());
}
static void make_norm(std::vector<uint16_t>& dst, int n, uint64_t seed, float norm_sigma) {
    dst.resize(n);
    if (norm_sigma > 0)
        kfo_fill_normal(dst.data(), n, seed, norm_sigma, 1.0f);
    else
        for (int i = 0; i < n; i++) dst[i] = 0x3f80;
}
extern "C" kfo_model* kfo_model_create(const kfo_model_config* cfg) {
    kfo_model* m = new kfo_model();
    m->c         = *cfg;
    const kfo_model_config& c = m->c;
    const int E = c.n_embd, F = c.n_ff, QD = c.n_head * c.head_dim, KD = c.n_kv_head * c.head_dim;
    m->w.resize(16 + 16 * (size_t)c.n_layer);
    auto S = [&](int id) { return kfo_tensor_seed(c.seed, id); };
    make_weight(m->w[0], c.vocab, E, c.embed_bits, c.embed_mode, c.group, S(0), c.sigma);
    make_norm(m->w[1], E, S(1), c.norm_sigma);
    if (!c.tie_embed) /* the quantizer keys select by tensor-name substring (G_Has_, GeQuant.cpp:1226): "embed_tokens" does not match lm_head.weight */
        make_weight(m->w[2], c.vocab, E, 16, 0, c.group, S(2), c.sigma);
    for (int l = 0; l < c.n_layer; l++) {
        const int b = 16 + 16 * l;
        make_norm(m->w[b + 0], E, S(b + 0), c.norm_sigma);
        make_weight(m->w[b + 1], QD, E, c.attn_bits, c.attn_mode, c.group, S(b + 1), c.sigma);
        make_weight(m->w[b + 2], KD, E, c.attn_bits, c.attn_mode, c.group, S(b + 2), c.sigma);
        make_weight(m->w[b + 3], KD, E, c.attn_bits, c.attn_mode, c.group, S(b + 3), c.sigma);
        make_norm(m->w[b + 4], c.head_dim, S(b + 4), c.norm_sigma);
        make_norm(m->w[b + 5], c.head_dim, S(b + 5), c.norm_sigma);
        make_weight(m->w[b + 6], E, QD, c.attn_bits, c.attn_mode, c.group, S(b + 6), c.sigma);
        make_norm(m->w[b + 7], E, S(b + 7), c.norm_sigma);
        make_weight(m->w[b + 8], F, E, c.mlp_bits, c.mlp_mode, c.group, S(b + 8), c.sigma);
        make_weight(m->w[b + 9], F, E, c.mlp_bits, c.mlp_mode, c.group, S(b + 9), c.sigma);
        make_weight(m->w[b + 10], E, F, c.mlp_bits, c.mlp_mode, c.group, S(b + 10), c.sigma);
    }
    m->kc.assign((size_t)c.n_layer * c.max_seq * KD, 0);
    m->vc.assign((size_t)c.n_layer * c.max_seq * KD, 0);
    return m;
}
extern "C" void kfo_model_destroy(kfo_model* m) { delete m; }
extern "C" void kfo_model_reset(kfo_model* m) {
    std::fill(m->kc.begin(), m->kc.end(), 0);
    std::fill(m->vc.begin(), m->vc.end(), 0);
}
extern "C" const uint16_t* kfo_model_weight(kfo_model* m, int id, size_t* n) {
    if (id < 0 || (size_t)id >= m->w.size())
        return nullptr;
    if (n)
        *n = m->w[id].size();
    return m->w[id].data();
}
extern "C" const uint16_t* kfo_model_kcache(kfo_model* m, int l) { return m->kc.data() + (size_t)l * m->c.max_seq * m->c.n_kv_head * m->c.head_dim; }
extern "C" const uint16_t* kfo_model_vcache(kfo_model* m, int l) { return m->vc.data() + (size_t)l * m->c.max_seq * m->c.n_kv_head * m->c.head_dim; }

/* one block: SelfAttention::cuInfer src/Device/CUDA/QKV.cu:617-706 then FFN::cuInfer src/Device/CUDA/NeuronFuse.cu:615-656 */
static void layer_forward(kfo_model* m, int l, int pos, std::vector<uint16_t>& x) {
    const kfo_model_config& c = m->c;
    const int E = c.n_embd, F = c.n_ff, hd = c.head_dim, QD = c.n_head * hd, KD = c.n_kv_head * hd, b = 16 + 16 * l;
    std::vector<uint16_t> h(E), q(QD), att(QD), o(E), g(F), u(F), s(F), d(E);
    uint16_t* krow = m->kc.data() + ((size_t)l * c.max_seq + pos) * KD; /* K.out/V.out alias the cache rows: TGraph.cpp:198-208 */
    uint16_t* vrow = m->vc.data() + ((size_t)l * c.max_seq + pos) * KD;
    kfo_rmsnorm(h.data(), x.data(), m->w[b + 0].data(), 1, E, c.rms_eps);
    kfo_linear(q.data(), m->w[b + 1].data(), h.data(), 1, QD, E);
    kfo_linear(krow, m->w[b + 2].data(), h.data(), 1, KD, E);
    kfo_linear(vrow, m->w[b + 3].data(), h.data(), 1, KD, E);
    /* ROPE::cuInfer src/Device/CUDA/kernel/rope.cu:645-672: QK-norm (per head) then rope; eps 1e-6 */
    kfo_rmsnorm(q.data(), q.data(), m->w[b + 4].data(), c.n_head, hd, 1e-6f);
    kfo_rmsnorm(krow, krow, m->w[b + 5].data(), c.n_kv_head, hd, 1e-6f);
    kfo_rope(q.data(), c.n_head, hd, pos, c.rope_theta);
    kfo_rope(krow, c.n_kv_head, hd, pos, c.rope_theta);
    kfo_attention_decode(att.data(), q.data(), kfo_model_kcache(m, l), kfo_model_vcache(m, l), pos, c.n_head, c.n_kv_head, hd, c.score_bf16);
    kfo_linear(o.data(), m->w[b + 6].data(), att.data(), 1, E, QD);
    kfo_add(x.data(), x.data(), o.data(), E);
    kfo_rmsnorm(h.data(), x.data(), m->w[b + 7].data(), 1, E, c.rms_eps);
    kfo_linear(g.data(), m->w[b + 8].data(), h.data(), 1, F, E);
    kfo_linear(u.data(), m->w[b + 9].data(), h.data(), 1, F, E);
    kfo_swiglu(s.data(), g.data(), u.data(), F);
    kfo_linear(d.data(), m->w[b + 10].data(), s.data(), 1, E, F);
    kfo_add(x.data(), x.data(), d.data(), E);
}
extern "C" int kfo_model_layer(kfo_model* m, int layer, int pos, uint16_t* x_inout) {
    if (layer < 0 || layer >= m->c.n_layer || pos < 0 || pos >= m->c.max_seq)
        return -1;
    std::vector<uint16_t> x(x_inout, x_inout + m->c.n_embd);
    layer_forward(m, layer, pos, x);
    memcpy(x_inout, x.data(), sizeof(uint16_t) * m->c.n_embd);
    return 0;
}
extern "C" int kfo_model_forward(kfo_model* m, int token, int pos, uint16_t* logits) {
    const kfo_model_config& c = m->c;
    if (token < 0 || token >= c.vocab || pos < 0 || pos >= c.max_seq)
        return -1;
    const int E = c.n_embd;
    std::vector<uint16_t> x(m->w[0].begin() + (size_t)token * E, m->w[0].begin() + (size_t)(token + 1) * E); /* TokenEmbed::cuInfer NeuronFuse.cu:176-207 */
    for (int l = 0; l < c.n_layer; l++) layer_forward(m, l, pos, x);
    if (logits) {
        std::vector<uint16_t> h(E);
        kfo_rmsnorm(h.data(), x.data(), m->w[1].data(), 1, E, c.rms_eps);                      /* final LayerNormal */
        kfo_linear(logits, (c.tie_embed ? m->w[0] : m->w[2]).data(), h.data(), 1, c.vocab, E); /* Head4Token::cuInfer_1 NeuronFuse.cu:842-862 */
    }
    return 0;
}
extern "C" int kfo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Sampler: GeneratOnPrompt::Sample, src/Manifold/GoPT.cpp:614-630 -- TopK :632-640 (TOPK_heap::Select :667-700), UpdateLogits :751-766,
 * TopP :729-748, Qu_FlipCoin :768-786, random_u32 / random_f32 :594-600.
 * selection 0: the top_k largest logits, ties to the lower index (what Select is meant to keep);
 * selection 1: what Select keeps as written: its std::priority_queue<int> orders the INDICES, so heap.top() is always the newest index and
 *              `isLarge(i, heap.top())` compares with the last pushed element -- the result is {0 .. k-2} plus the first maximum of the rest.
 * Candidates are then ordered by (logit descending, index ascending); the reference's std::sort on `a > b` leaves the order of equal
 * logits unspecified -- this is one valid outcome of it.
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t samp_key(uint16_t b) { return (b & 0x8000u) ? (uint32_t)(uint16_t)~b : (uint32_t)(b | 0x8000u); }
extern "C" int kfo_sample(const uint16_t* logits, int vocab, float temperature, int top_k, float top_p, uint64_t* rng_state, int selection,
                          int* n_pick_out) {
    if (temperature == 0.0f || top_k == 1) { /* sample_argmax :602-612 */
        int best = 0;
        for (int i = 1; i < vocab; i++)
            if (kfo_bf16_to_f32(logits[i]) > kfo_bf16_to_f32(logits[best])) best = i;
        return best;
    }
    if (top_k <= 0 || top_k > vocab) top_k = vocab;
    std::vector<int> picks;
    if (selection == 1) {
        for (int i = 0; i < top_k - 1; i++) picks.push_back(i);
        int best = top_k - 1;
        for (int i = top_k; i < vocab; i++)
            if (kfo_bf16_to_f32(logits[i]) > kfo_bf16_to_f32(logits[best])) best = i;
        picks.push_back(best);
    } else {
        std::vector<int> idx(vocab);
        for (int i = 0; i < vocab; i++) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return samp_key(logits[a]) > samp_key(logits[b]); });
        picks.assign(idx.begin(), idx.begin() + top_k);
    }
    std::stable_sort(picks.begin(), picks.end(), [&](int a, int b) {
        const float va = kfo_bf16_to_f32(logits[a]), vb = kfo_bf16_to_f32(logits[b]);
        return va > vb || (va == vb && a < b);
    });
    std::vector<float> p(top_k);
    const float mx = kfo_bf16_to_f32(logits[picks[0]]);
    float sum      = 0.f;
    for (int i = 0; i < top_k; i++) {
        p[i] = expf((kfo_bf16_to_f32(logits[picks[i]]) - mx) / temperature);
        sum += p[i];
    }
    for (int i = 0; i < top_k; i++) p[i] /= sum;
    int n_pick = top_k;
    if (top_p < 1.0f) {
        float cum = 0.f;
        int last  = top_k - 1;
        for (int i = 0; i < top_k; i++) {
            cum += p[i];
            if (cum > top_p) {
                last = i;
                break;
            }
        }
        n_pick = last + 1;
    }
    if (n_pick_out) *n_pick_out = n_pick;
    float psum = 0.f;
    for (int i = 0; i < n_pick; i++) psum += p[i];
    uint64_t s = *rng_state;
    s ^= s >> 12, s ^= s << 25, s ^= s >> 27;
    *rng_state       = s;
    const uint32_t r = (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
    const float coin = (float)(r >> 8) / 16777216.0f * psum;
    int q            = picks[n_pick - 1];
    float cdf        = 0.f;
    for (int i = 0; i < n_pick; i++) {
        cdf += p[i];
        if (coin < cdf) {
            q = picks[i];
            break;
        }
    }
    return q;
}
