// awq.cu -- the vendor AWQ layout as the reference reads it (typNUMBER::Q4 under QUANT_MODE::AWQ): CU_Q42X_awq, reference
// src/Device/CUDA/kernel/quantizer.cu:132-156, CU_I2Q4_unpack + AWQ_REVERSE_ORDER src/Device/CUDA/kernel/packedN.cuh:109-116; the matmul uses
// the tensor with transA = 0 (SLP::Forw, NeuronFuse.cu:305-381), i.e. it is STORED [in_features][out_features]:
//     qweight int32 [IC][OC / 8]   element 8c + k of a row in nibble AWQ_REVERSE_ORDER[k] = {0,4,1,5,2,6,3,7}[k] of word c
//     qzeros  int32 [IC / 128][OC / 8]   same nibble order          scales fp16 [IC / 128][OC]
//     w[ic][oc] = bf16( float(q - z) * float(scale) )
// kf_tensor_desc for it: type KF_T_AWQ4, rows = OC, cols = IC, data_dev = qweight, zero_dev = qzeros, step_dev = scales, group = 128.
// AWQ tensors come from vendor checkpoints (the reference has no AWQ quantiser, and neither has this library): this is the device-level
// entry (dequant + matmul); the model runtime's loader for such checkpoints is a 'next' row.
//
// Matmul: up to 8 tokens a coalesced column-walking GEMV (a thread owns one int32 word = 8 output columns and strides over the input rows;
// k-slices reduced in a fixed order by a second small launch), more tokens dequantise into [OC][IC] bf16 + the tcgen05 GEMM.  Functional
// path, not tuned.
#include <cuda_fp16.h>

#include "kf_common.cuh"

namespace {
__device__ __forceinline__ int awq_nibble(uint32_t word, int k) {  // element k of the 8 in a word
    const int order = ((k & 1) << 2) | (k >> 1);                   // AWQ_REVERSE_ORDER = {0,4,1,5,2,6,3,7}
    return (int)((word >> (order * 4)) & 0xFu);
}
__device__ __forceinline__ float awq_weight(int q, int z, uint16_t scale_f16) {
    return bf16_bits_to_f32(f32_to_bf16_bits((float)(q - z) * __half2float(__ushort_as_half(scale_f16))));
}

// out[oc][ic] (the library's [rows][cols] view) or out[ic][oc] (the reference's GetDataX order): grid (OC / 8 / 32, IC / 32)
__global__ void __launch_bounds__(256) kf_awq_dequant_kernel(const uint32_t* __restrict__ qw, const uint32_t* __restrict__ qz, const uint16_t* __restrict__ sc,
                                                             int IC, int OC, uint16_t* __restrict__ out, int transposed) {
    __shared__ uint16_t tile[32][256 + 2];  // [ic][oc] of a 32 x 256 block
    const int wl = threadIdx.x & 31, il = threadIdx.x >> 5;  // word lane, ic lane (8)
    const int word = blockIdx.x * 32 + wl, W8 = OC / 8;
    const int ic0  = blockIdx.y * 32;
    for (int r = il; r < 32; r += 8) {
        const int ic = ic0 + r;
        if (word < W8 && ic < IC) {
            const uint32_t q = qw[(size_t)ic * W8 + word], z = qz[(size_t)(ic / 128) * W8 + word];
            const uint4 s8   = *reinterpret_cast<const uint4*>(sc + (size_t)(ic / 128) * OC + (size_t)word * 8);
            const uint32_t sw[4] = {s8.x, s8.y, s8.z, s8.w};
#pragma unroll
            for (int k = 0; k < 8; k++)
                tile[r][wl * 8 + k] = f32_to_bf16_bits((float)(awq_nibble(q, k) - awq_nibble(z, k)) *
                                                       __half2float(__ushort_as_half((uint16_t)(sw[k >> 1] >> ((k & 1) * 16)))));
        }
    }
    __syncthreads();
    if (!transposed) {  // [ic][oc]
        for (int r = il; r < 32; r += 8) {
            const int ic = ic0 + r;
            if (word < W8 && ic < IC)
#pragma unroll
                for (int k = 0; k < 8; k++) out[(size_t)ic * OC + (size_t)word * 8 + k] = tile[r][wl * 8 + k];
        }
    } else {  // [oc][ic]: a warp writes 32 consecutive ic of one oc
        for (int c = il; c < 256; c += 8) {
            const int oc = blockIdx.x * 256 + c, ic = ic0 + wl;
            if (oc < OC && ic < IC) out[(size_t)oc * IC + ic] = tile[wl][c];
        }
    }
}

// partial[split][m][oc] = sum over the split's input rows; block = 32 words x 8 ic lanes; grid (OC / 256, S)
template <int MT>
__global__ void __launch_bounds__(256) kf_awq_gemv_kernel(float* __restrict__ partial, const uint32_t* __restrict__ qw, const uint32_t* __restrict__ qz,
                                                          const uint16_t* __restrict__ sc, const uint16_t* __restrict__ x, int IC, int OC, int M, int S) {
    __shared__ float red[8][32][9];
    const int wl = threadIdx.x & 31, il = threadIdx.x >> 5;
    const int word = blockIdx.x * 32 + wl, W8 = OC / 8;
    const int groups = IC / 128, g0 = (int)(((long long)blockIdx.y * groups) / S), g1 = (int)(((long long)(blockIdx.y + 1) * groups) / S);
    kf_grid_dependency_wait();
    float acc[MT][8];
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[m][k] = 0.f;
    if (word < W8) {
        for (int g = g0; g < g1; g++) {
            const uint32_t z = __ldg(qz + (size_t)g * W8 + word);
            const uint4 s8   = __ldg(reinterpret_cast<const uint4*>(sc + (size_t)g * OC + (size_t)word * 8));
            const uint32_t sw[4] = {s8.x, s8.y, s8.z, s8.w};
            int zk[8];
            uint16_t sk[8];
#pragma unroll
            for (int k = 0; k < 8; k++) zk[k] = awq_nibble(z, k), sk[k] = (uint16_t)(sw[k >> 1] >> ((k & 1) * 16));
            for (int r = il; r < 128; r += 8) {
                const int ic     = g * 128 + r;
                const uint32_t q = __ldg(qw + (size_t)ic * W8 + word);
                float wk[8];
#pragma unroll
                for (int k = 0; k < 8; k++) wk[k] = awq_weight(awq_nibble(q, k), zk[k], sk[k]);
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    if (m >= M) break;
                    const float xv = bf16_bits_to_f32(__ldg(x + (size_t)m * IC + ic));
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[m][k] = fmaf(wk[k], xv, acc[m][k]);
                }
            }
        }
    }
    for (int m = 0; m < MT && m < M; m++) {  // the 8 ic lanes of a word, added in lane order
#pragma unroll
        for (int k = 0; k < 8; k++) red[il][wl][k] = acc[m][k];
        __syncthreads();
        if (il == 0 && word < W8) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float s = 0.f;
                for (int l = 0; l < 8; l++) s += red[l][wl][k];
                partial[((size_t)blockIdx.y * M + m) * OC + (size_t)word * 8 + k] = s;
            }
        }
        __syncthreads();
    }
}
// y[m][oc] = epilogue(sum over splits, in split order)
__global__ void __launch_bounds__(256) kf_awq_reduce_kernel(void* __restrict__ y, const float* __restrict__ partial, const uint16_t* __restrict__ residual,
                                                            size_t n, int S, int epi) {
    kf_grid_dependency_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < S; p++) s += partial[(size_t)p * n + i];
    if (epi == KF_EPI_F32) {
        reinterpret_cast<float*>(y)[i] = s;
        return;
    }
    uint16_t v = f32_to_bf16_bits(s);
    if (epi == KF_EPI_RESIDUAL) v = f32_to_bf16_bits(bf16_bits_to_f32(residual[i]) + bf16_bits_to_f32(v));
    reinterpret_cast<uint16_t*>(y)[i] = v;
}
}  // namespace

static bool awq_ok(const kf_tensor_desc* w) {
    return w->data_dev && w->zero_dev && w->step_dev && w->rows % 8 == 0 && w->cols % 128 == 0 && w->group == 128 &&
           (((uintptr_t)w->data_dev | (uintptr_t)w->zero_dev | (uintptr_t)w->step_dev) & 15) == 0;
}
// transposed = 1: out[rows = OC][cols = IC] (the library's view of a weight); 0: [IC][OC], the order of the reference's GetDataX
int kf_awq_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out, int transposed) {
    KF_REQUIRE(ctx, awq_ok(w), "AWQ tensor: qweight / qzeros / scales, out_features % 8 == 0, in_features % 128 == 0, group 128, 16-byte aligned");
    const int OC = w->rows, IC = w->cols;
    dim3 grid((OC / 8 + 31) / 32, (IC + 31) / 32);
    kf_awq_dequant_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t*)w->data_dev, (const uint32_t*)w->zero_dev, (const uint16_t*)w->step_dev, IC, OC,
                                                          (uint16_t*)out, transposed);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
// M <= 8
int kf_awq_gemv(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    KF_REQUIRE(ctx, awq_ok(w) && M >= 1 && M <= 8, "AWQ GEMV: a well-formed AWQ tensor, up to 8 tokens");
    KF_REQUIRE(ctx, epilogue == KF_EPI_NONE || epilogue == KF_EPI_F32 || (epilogue == KF_EPI_RESIDUAL && residual), "epilogue");
    const int OC = w->rows, IC = w->cols, cols_ctas = (OC / 8 + 31) / 32;
    int S = std::max(1, std::min(IC / 128, (2 * ctx->sm_count + cols_ctas - 1) / cols_ctas));  // ~two waves of CTAs
    const size_t n = (size_t)M * OC;
    int rc = kf_ensure_buf(ctx, &ctx->awq_ws, &ctx->awq_ws_bytes, (size_t)S * n * 4);
    if (rc) return rc;
    float* partial = (float*)ctx->awq_ws;
    const dim3 grid(cols_ctas, S);
#define KF_AWQ_GO(MT)                                                                                                                  \
    KF_CUDA(ctx, kf_launch_pdl(ctx, kf_awq_gemv_kernel<MT>, grid, dim3(256), 0, partial, (const uint32_t*)w->data_dev, (const uint32_t*)w->zero_dev, \
                               (const uint16_t*)w->step_dev, (const uint16_t*)x, IC, OC, M, S))
    if (M == 1)
        KF_AWQ_GO(1);
    else if (M == 2)
        KF_AWQ_GO(2);
    else if (M <= 4)
        KF_AWQ_GO(4);
    else
        KF_AWQ_GO(8);
#undef KF_AWQ_GO
    KF_LAUNCH_CHECK(ctx);
    KF_CUDA(ctx, kf_launch_pdl(ctx, kf_awq_reduce_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, y, (const float*)partial,
                               (const uint16_t*)residual, n, S, epilogue));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
