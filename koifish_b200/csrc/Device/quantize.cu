// quantize.cu -- synthetic weight fill, quantise-at-load (pack to 128-bit words) and the GetDataX test hook.
//
//  kf_fill_normal : replaces CU_disti_normal in huTensor::InitParam (reference src/Device/CUDA/huTensor.cu:199-210) with a
//                   counter-based generator that a CPU can reproduce bit-for-bit.
//  kf_quantize    : GeQuant::LowBit_worker / RTN_x / YinYang (reference src/Tensor/GeQuant.cpp:830-905, 428-533, 536-628) on the
//                   device.  It reproduces the CPU packer's float arithmetic operation by operation (IEEE division/rounding
//                   intrinsics, sequential double sum per group), so the bytes equal what the reference's load path writes.
//  kf_dequant     : GTensor::GetDataX (reference src/Device/CUDA/kernel/quantizer.cu:249-392 -> CU_Q128toX_, T.cu:245-294).
#include <algorithm>

#include "kf_common.cuh"

// ------------------------------------------------------------------------------------------------ fill
__device__ __forceinline__ uint64_t kf_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) kf_fill_normal_kernel(uint16_t* __restrict__ out, size_t n, uint64_t seed_mul, float scale, float mean) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = kf_mix64(seed_mul + i);
        int s      = (int)(h & 0xffff) + (int)((h >> 16) & 0xffff) + (int)((h >> 32) & 0xffff) + (int)((h >> 48) & 0xffff);
        float z    = (float)(s - 131070);
        out[i]     = f32_to_bf16_bits(__fmaf_rn(z, scale, mean));
    }
}
__global__ void __launch_bounds__(256) kf_fill_normal_2d_kernel(uint16_t* __restrict__ out, int rows, int cols, size_t ld, size_t row0, size_t col0,
                                                                uint64_t seed_mul, float scale, float mean) {
    const size_t n = (size_t)rows * cols, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t r = i / cols, c = i - r * cols;
        uint64_t h = kf_mix64(seed_mul + (row0 + r) * ld + col0 + c);
        int s      = (int)(h & 0xffff) + (int)((h >> 16) & 0xffff) + (int)((h >> 32) & 0xffff) + (int)((h >> 48) & 0xffff);
        out[i]     = f32_to_bf16_bits(__fmaf_rn((float)(s - 131070), scale, mean));
    }
}
extern "C" int kf_fill_normal_2d(kf_ctx* ctx, void* out, int rows, int cols, size_t ld, size_t row0, size_t col0, uint64_t seed, float sigma,
                                 float mean) {
    if (!ctx || !out || rows < 0 || cols < 0) return KF_ERR_BAD_ARG;
    const size_t n = (size_t)rows * cols;
    if (n == 0) return KF_OK;
    size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 32);
    kf_fill_normal_2d_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((uint16_t*)out, rows, cols, ld, row0, col0, seed * 0xD1342543DE82EF95ull,
                                                                        sigma / 37837.227f, mean);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
extern "C" int kf_fill_normal(kf_ctx* ctx, void* out, size_t n, uint64_t seed, float sigma, float mean) {
    if (!ctx || !out)
        return KF_ERR_BAD_ARG;
    if (n == 0)
        return KF_OK;
    const float scale = sigma / 37837.227f;  // std of the sum of four U{0..65535}
    size_t blocks     = (n + 255) / 256;
    if (blocks > (size_t)ctx->sm_count * 32)
        blocks = (size_t)ctx->sm_count * 32;
    kf_fill_normal_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((uint16_t*)out, n, seed * 0xD1342543DE82EF95ull, scale, mean);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// ------------------------------------------------------------------------------------------------ quantise
// One thread per group, sequential in the order the CPU packer walks it (so the double-precision energy sum is the same).
template <int BITS>
__global__ void __launch_bounds__(128) kf_quantize_kernel(const uint16_t* __restrict__ w, size_t nG, int group, int mode, int qMin, int qMax,
                                                           int qBias, uint8_t* __restrict__ data, uint16_t* __restrict__ gZero,
                                                           uint16_t* __restrict__ gStep) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nG)
        return;
    constexpr int PER = 128 / BITS, HALF = PER / 2;
    const uint16_t* dat = w + g * group;
    float vmax = -3.402823466e+38f, vmin = 3.402823466e+38f;
    double vSum = 0.0;
    for (int i = 0; i < group; i++) {
        float a = bf16_bits_to_f32(dat[i]);
        vmax = fmaxf(vmax, a), vmin = fminf(vmin, a);
        if (BITS == 1)
            vSum += a < 0.0f ? 0.0 : (double)__fmul_rn(a, a);  // GeQuant.cpp:573
        else
            vSum += (double)fabsf(a);                           // GeQuant.cpp:461
    }
    float step, zero;
    if (BITS == 1) {
        float vMean = (float)sqrt(vSum / (double)group);
        step = fmaxf(1e-5f, vMean), zero = 0.f;
    } else {
        float vMean = (float)(vSum / (double)group);
        step = __fdiv_rn(__fsub_rn(vmax, vmin), (float)(qMax - qMin)), zero = -vmin;
        if (mode == KF_Q_YYANG) {
            step = fmaxf(1e-5f, vMean), zero = 0.f;
        } else if (mode == KF_Q_RTN_SYM) {
            step = __fdiv_rn(fmaxf(fabsf(vmax), fabsf(vmin)), (float)qMax), zero = 0.f;
        }
    }
    gZero[g] = f32_to_bf16_bits(zero), gStep[g] = f32_to_bf16_bits(step);
    const bool clampq = (mode == KF_Q_YYANG) || BITS == 1;
    uint8_t* quanti   = data + g * (size_t)group * BITS / 8;
    for (int wd = 0; wd < group / PER; wd++) {
        unsigned long long high = 0, low = 0;
        for (int pos = 0; pos < PER; pos++) {
            float a = bf16_bits_to_f32(dat[wd * PER + pos]);
            int qid = 0;
            if (step != 0.0f)
                qid = (int)roundf(__fdiv_rn(__fadd_rn(a, zero), step));
            if (clampq)
                qid = min(max(qid, qMin), qMax);
            unsigned long long code = (unsigned long long)((qid + qBias) & ((1 << BITS) - 1));
            if (pos < HALF)
                high |= code << (64 - BITS * (pos + 1));
            else
                low |= code << (64 - BITS * (pos - HALF + 1));
        }
        ulonglong2 v;
        v.x = low, v.y = high;  // struct Packed128 { low, high } (PackedQ.hpp:28-31)
        *reinterpret_cast<ulonglong2*>(quanti + 16 * wd) = v;
    }
}
__global__ void __launch_bounds__(256) kf_f8e5m2_encode_kernel(const uint16_t* __restrict__ w, size_t n, uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        __half h = __float2half_rn(bf16_bits_to_f32(w[i]));  // packedN.cuh:87
        out[i]   = (uint8_t)(__half_as_ushort(h) >> 8);       // keep the high byte
    }
}

extern "C" size_t kf_quant_data_bytes(int rows, int cols, int type) { return (size_t)rows * cols * kf_type_bits(type) / 8; }
extern "C" size_t kf_quant_gama_bytes(int rows, int cols, int type, int group) {
    if (type == KF_T_NF4) return 2 * ((size_t)rows + cols + 16 * (size_t)rows);  // szGama of RT_NormalF, GeQuant.cpp:744
    if (!kf_type_packed(type) || group <= 0)
        return 0;
    return 2 * ((size_t)rows + cols + 2 * ((size_t)rows * cols / group));  // szGama, GeQuant.cpp:518
}

static int qrange_of(int bits, int mode, int* qMin, int* qMax, int* qBias) {
    if (mode == KF_Q_YYANG) {
        if (bits == 2)
            *qMax = 1, *qMin = -1, *qBias = 1;
        else if (bits == 1)
            *qMax = 1, *qMin = 0, *qBias = 0;
        else
            return -1;
    } else if (mode == KF_Q_RTN_SYM) {
        *qMin = -(1 << (bits - 1)), *qMax = (1 << (bits - 1)) - 1, *qBias = -*qMin;
    } else {
        *qMin = 0, *qMax = (1 << bits) - 1, *qBias = 0;
    }
    return 0;
}

extern "C" int kf_quantize(kf_ctx* ctx, const void* w, int rows, int cols, int type, int group, int mode, void* data, void* gama, int* qbias_out) {
    if (!ctx || !w || !data)
        return KF_ERR_BAD_ARG;
    const size_t n = (size_t)rows * cols;
    if (type == KF_T_BF16) {
        KF_CUDA(ctx, cudaMemcpyAsync(data, w, n * 2, cudaMemcpyDeviceToDevice, ctx->stream));
        if (qbias_out)
            *qbias_out = 0;
        return KF_OK;
    }
    if (type == KF_T_F8E5M2) {
        size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 32);
        kf_f8e5m2_encode_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint16_t*)w, n, (uint8_t*)data);
        KF_LAUNCH_CHECK(ctx);
        if (qbias_out)
            *qbias_out = 0;
        return KF_OK;
    }
    KF_REQUIRE(ctx, type != KF_T_AWQ4, "AWQ tensors come packed from vendor checkpoints: there is no AWQ quantiser (nor has the reference one)");
    if (type == KF_T_NF4) {
        if (qbias_out) *qbias_out = 0;
        return kf_nf4_quantize(ctx, w, rows, cols, data, gama);
    }
    KF_REQUIRE(ctx, kf_type_packed(type) && gama, "packed type needs a gama buffer");
    const int bits = kf_type_bits(type);
    if (type == KF_T_SIGN || type == KF_T_BINARY)
        KF_REQUIRE(ctx, mode == KF_Q_YYANG, "T_SIGN / T_BINARY are produced by the yyang quantiser only");
    if (bits == 1)
        KF_REQUIRE(ctx, mode == KF_Q_YYANG, "1-bit goes through YinYang (GeQuant.cpp:909)");
    int qMin, qMax, qBias;
    KF_REQUIRE(ctx, qrange_of(bits, mode, &qMin, &qMax, &qBias) == 0, "bits/mode");
    KF_REQUIRE(ctx, group > 0 && n % group == 0 && group % (128 / bits) == 0 && cols % group == 0,
               "group must divide the row and hold whole 128-bit words (GeQuant.cpp:438)");
    const size_t nG = n / group;
    uint16_t* g0    = (uint16_t*)gama;
    KF_CUDA(ctx, cudaMemsetAsync(g0, 0, 2 * ((size_t)rows + cols), ctx->stream));  // R/C scales unused (NO_NORMAL)
    uint16_t* gZero = g0 + rows + cols;
    uint16_t* gStep = gZero + nG;
    unsigned blocks = (unsigned)((nG + 127) / 128);
    if (bits == 4)
        kf_quantize_kernel<4><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)w, nG, group, mode, qMin, qMax, qBias, (uint8_t*)data, gZero, gStep);
    else if (bits == 2)
        kf_quantize_kernel<2><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)w, nG, group, mode, qMin, qMax, qBias, (uint8_t*)data, gZero, gStep);
    else
        kf_quantize_kernel<1><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)w, nG, group, mode, qMin, qMax, qBias, (uint8_t*)data, gZero, gStep);
    KF_LAUNCH_CHECK(ctx);
    if (qbias_out)
        *qbias_out = qBias;
    return KF_OK;
}

// ------------------------------------------------------------------------------------------------ dequant (GetDataX test hook)
// One thread per 128-bit word.  Arithmetic exactly as CU_Q128toX_ (T.cu:274); FMA: see kf_deq_scalar (kf_common.cuh).
template <int BITS, bool FMA>
__global__ void __launch_bounds__(256) kf_dequant_kernel(const uint4* __restrict__ words, size_t nWords, int group, int qBias,
                                                          const uint16_t* __restrict__ gZero, const uint16_t* __restrict__ gStep,
                                                          uint16_t* __restrict__ out) {
    const size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nWords)
        return;
    constexpr int PER = 128 / BITS, HALF = PER / 2;
    const uint4 q   = words[wi];
    const unsigned long long low = ((unsigned long long)q.y << 32) | q.x, high = ((unsigned long long)q.w << 32) | q.z;
    const size_t e0 = wi * PER;
    const size_t g  = e0 / group;
    const ushort2 zero_step = make_ushort2(gZero[g], gStep[g]);
#pragma unroll 8
    for (int j = 0; j < PER; j++) {
        const unsigned long long src = j < HALF ? high : low;
        const int jj   = j < HALF ? j : j - HALF;
        const int code = (int)((src >> (64 - BITS * (jj + 1))) & ((1u << BITS) - 1));
        out[e0 + j]    = kf_deq_scalar<FMA>(code - qBias, zero_step.y, zero_step.x);
    }
}
__global__ void __launch_bounds__(256) kf_f8e5m2_decode_kernel(const uint8_t* __restrict__ in, size_t n, uint16_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = f32_to_bf16_bits(__half2float(__ushort_as_half((unsigned short)((unsigned short)in[i] << 8))));
}
extern "C" int kf_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out) {
    if (!ctx || !w || !out || !w->data_dev)
        return KF_ERR_BAD_ARG;
    const size_t n = (size_t)w->rows * w->cols;
    if (w->type == KF_T_BF16) {
        KF_CUDA(ctx, cudaMemcpyAsync(out, w->data_dev, n * 2, cudaMemcpyDeviceToDevice, ctx->stream));
        return KF_OK;
    }
    if (w->type == KF_T_F8E5M2) {
        size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 32);
        kf_f8e5m2_decode_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint8_t*)w->data_dev, n, (uint16_t*)out);
        KF_LAUNCH_CHECK(ctx);
        return KF_OK;
    }
    if (w->type == KF_T_NF4) return kf_nf4_dequant(ctx, w, out);
    if (w->type == KF_T_AWQ4) return kf_awq_dequant(ctx, w, out, 1);  // [rows = out_features][cols = in_features], like every other type
    KF_REQUIRE(ctx, kf_type_packed(w->type) && kf_has_gama(*w) && w->group > 0, "packed tensor needs gama + group");
    const int bits = kf_type_bits(w->type), per = 128 / bits;
    KF_REQUIRE(ctx, n % w->group == 0 && w->group % per == 0, "group / word alignment");
    const size_t nWords = n / per;
    unsigned blocks     = (unsigned)((nWords + 255) / 256);
    const uint16_t *gz = kf_gama_zero(*w), *gs = kf_gama_step(*w);
#define KF_DEQ_LAUNCH(B)                                                                                                                      \
    (ctx->deq_fma ? kf_dequant_kernel<B, true><<<blocks, 256, 0, ctx->stream>>>((const uint4*)w->data_dev, nWords, w->group, w->qbias, gz, gs,  \
                                                                                (uint16_t*)out)                                                 \
                  : kf_dequant_kernel<B, false><<<blocks, 256, 0, ctx->stream>>>((const uint4*)w->data_dev, nWords, w->group, w->qbias, gz, gs, \
                                                                                 (uint16_t*)out))
    if (bits == 4)
        KF_DEQ_LAUNCH(4);
    else if (bits == 2)
        KF_DEQ_LAUNCH(2);
    else
        KF_DEQ_LAUNCH(1);
#undef KF_DEQ_LAUNCH
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
