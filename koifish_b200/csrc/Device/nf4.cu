// nf4.cu -- the reference's NormalFloat format: QUANT_MODE::RTNf, what a quantizer entry {"bits": 4} WITHOUT a quant_method selects
// (QUANT_CARD::Init4Neuron, src/Tensor/GeQuant.cpp:1270-1280).
//
//   quantise  GeQuant::RT_NormalF / _row_lut (GeQuant.cpp:696-748) with Distri_PIPE::Prepare (symmetric case, :674-681) and X2NormalF (:684-700):
//             per ROW  abs_max = max |w| ; scale = abs_max > 0 ? 1 / abs_max : 1 ; codebook[i] = NF4_table[i] / scale (fp32) ;
//             code = the nearest codebook entry (first minimum of |w - codebook[i]|, fp32) ; gama LUT = bf16(codebook)
//   storage   data: an MSB-first bit stream, element i at bit 4 i (BIT_SET_k, src/Utils/CLI_params.cpp:2177-2191): byte j holds element 2j in
//             its HIGH nibble and 2j + 1 in its low one ; gama: bf16 [rows R_SCALE][cols C_SCALE][rows][16] (szGama, GeQuant.cpp:744)
//   dequant   CU_Q42X_NF4 (src/Device/CUDA/kernel/quantizer.cu:612-654) with rc_normal = 0:  w = lut[row][code]
//
// Matmul: up to 8 tokens take a warp-per-row GEMV with a per-row 256-entry (byte -> two weights) table in shared memory; more tokens dequantise
// into a context scratch and take the bf16 tcgen05 GEMM, as the reference does for every format (GTensor::GetDataX + cuBLASLt).  This is the
// functional path of the format, not a tuned one: see DESIGN.md.
#include "kf_common.cuh"

namespace {
// NF4_LUT::table, src/g_float.hpp:543-558 (the QLoRA NormalFloat4 quantiles)
__constant__ float kNF4[16] = {-1.0f,
                               -0.6961928009986877f,
                               -0.5250730514526367f,
                               -0.39491748809814453f,
                               -0.28444138169288635f,
                               -0.18477343022823334f,
                               -0.09105003625154495f,
                               0.0f,
                               0.07958029955625534f,
                               0.16093020141124725f,
                               0.24611230194568634f,
                               0.33791524171829224f,
                               0.44070982933044434f,
                               0.5626170039176941f,
                               0.7229568362236023f,
                               1.0f};

// one CTA per row
__global__ void __launch_bounds__(256) kf_nf4_quantize_kernel(const uint16_t* __restrict__ w, int rows, int cols, uint8_t* __restrict__ data,
                                                              uint16_t* __restrict__ lut) {
    __shared__ float s_red[32];
    __shared__ float s_cb[16];
    const int row = blockIdx.x, tid = threadIdx.x;
    const uint16_t* wr = w + (size_t)row * cols;
    float amax = 0.f;
    for (int i = tid; i < cols; i += blockDim.x) amax = fmaxf(amax, fabsf(bf16_bits_to_f32(wr[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = amax;
    __syncthreads();
    if (tid < 16) {
        float m = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) m = fmaxf(m, s_red[i]);
        const float scale = m > 0.f ? __fdiv_rn(1.0f, m) : 1.0f;
        const float cb    = __fdiv_rn(kNF4[tid], scale);
        s_cb[tid]         = cb;
        lut[(size_t)row * 16 + tid] = f32_to_bf16_bits(cb);
    }
    __syncthreads();
    float cb[16];
#pragma unroll
    for (int i = 0; i < 16; i++) cb[i] = s_cb[i];
    uint8_t* dr = data + (size_t)row * cols / 2;
    for (int j = tid; j < cols / 2; j += blockDim.x) {
        int code[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float a = bf16_bits_to_f32(wr[2 * j + h]);
            float best    = 3.402823466e+38f;
            int bi        = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float d = fabsf(a - cb[i]);
                if (d < best) best = d, bi = i;
            }
            code[h] = bi;
        }
        dr[j] = (uint8_t)((code[0] << 4) | code[1]);
    }
}

__global__ void __launch_bounds__(256) kf_nf4_dequant_kernel(const uint8_t* __restrict__ data, const uint16_t* __restrict__ lut, int cols,
                                                             uint16_t* __restrict__ out) {
    __shared__ uint16_t s_lut[16];
    const int row = blockIdx.x;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = lut[(size_t)row * 16 + threadIdx.x];
    __syncthreads();
    const uint8_t* dr = data + (size_t)row * cols / 2;
    uint32_t* orow    = reinterpret_cast<uint32_t*>(out + (size_t)row * cols);
    for (int j = threadIdx.x; j < cols / 2; j += blockDim.x) {
        const uint8_t b = dr[j];
        orow[j]         = (uint32_t)s_lut[b >> 4] | ((uint32_t)s_lut[b & 15] << 16);
    }
}
// embedding rows: grid (cols / 512, M)
__global__ void __launch_bounds__(256) kf_nf4_embed_kernel(uint16_t* __restrict__ out, const uint8_t* __restrict__ data, const uint16_t* __restrict__ lut,
                                                           const int32_t* __restrict__ tokens, int rows, int cols) {
    const int m = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    int tok = tokens[m];
    tok     = tok < 0 ? 0 : (tok >= rows ? rows - 1 : tok);
    if (j >= cols / 2) return;
    const uint8_t b      = data[(size_t)tok * cols / 2 + j];
    const uint16_t* lrow = lut + (size_t)tok * 16;
    reinterpret_cast<uint32_t*>(out + (size_t)m * cols)[j] = (uint32_t)lrow[b >> 4] | ((uint32_t)lrow[b & 15] << 16);
}

// y[m][row] = sum_k lut[row][code(row, k)] * x[m][k], fp32 accumulate.  One warp per row, 8 rows per CTA; a lane takes 16 bytes (32 codes).
// epi: 0 bf16 ; 1 bf16(residual + bf16(acc)) ; 4 fp32
template <int MT>
__global__ void __launch_bounds__(256) kf_nf4_gemv_kernel(void* __restrict__ y, const uint8_t* __restrict__ data, const uint16_t* __restrict__ lut,
                                                          const uint16_t* __restrict__ x, const uint16_t* __restrict__ residual, int rows, int K, int M,
                                                          int epi) {
    __shared__ float2 s_pair[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row  = blockIdx.x * 8 + warp;
    kf_grid_dependency_wait();
    if (row >= rows) return;
    {
        const uint16_t* lr = lut + (size_t)row * 16;
        for (int b = lane; b < 256; b += 32) s_pair[warp][b] = make_float2(bf16_bits_to_f32(lr[b >> 4]), bf16_bits_to_f32(lr[b & 15]));
    }
    __syncwarp();
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; m++) acc[m] = 0.f;
    const uint8_t* dr = data + (size_t)row * K / 2;
    for (int k0 = lane * 32; k0 < K; k0 += 1024) {
        const uint4 wv      = __ldg(reinterpret_cast<const uint4*>(dr + k0 / 2));
        const uint32_t wq[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int m = 0; m < MT; m++) {
            if (m >= M) break;
            const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)m * K + k0);
            float a          = acc[m];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 xv       = __ldg(xp + q);  // 8 activations: the two bytes q*... of register q>>... (see below)
                const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {  // byte 4q + j of the 16 holds elements 8q + 2j (high nibble) and 8q + 2j + 1
                    const uint32_t byte = (wq[q] >> (8 * j)) & 0xffu;
                    const float2 wp     = s_pair[warp][byte];
                    a = fmaf(wp.x, bf16lo(xs[j]), a);
                    a = fmaf(wp.y, bf16hi(xs[j]), a);
                }
            }
            acc[m] = a;
        }
    }
#pragma unroll
    for (int m = 0; m < MT; m++) {
        float a = acc[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && m < M) {
            const size_t at = (size_t)m * rows + row;
            if (epi == 4) {
                reinterpret_cast<float*>(y)[at] = a;
            } else {
                uint16_t v = f32_to_bf16_bits(a);
                if (epi == 1) v = f32_to_bf16_bits(bf16_bits_to_f32(residual[at]) + bf16_bits_to_f32(v));
                reinterpret_cast<uint16_t*>(y)[at] = v;
            }
        }
    }
}
}  // namespace

static const uint16_t* nf4_lut(const kf_tensor_desc& w) { return (const uint16_t*)w.gama_dev + w.rows + w.cols; }

int kf_nf4_quantize(kf_ctx* ctx, const void* w, int rows, int cols, void* data, void* gama) {
    KF_REQUIRE(ctx, gama && cols % 32 == 0 && rows >= 1, "NF4: a gama buffer, cols a multiple of 32");
    uint16_t* g0 = (uint16_t*)gama;
    KF_CUDA(ctx, cudaMemsetAsync(g0, 0, 2 * ((size_t)rows + cols), ctx->stream));  // R/C scales unused (NO_NORMAL)
    kf_nf4_quantize_kernel<<<rows, 256, 0, ctx->stream>>>((const uint16_t*)w, rows, cols, (uint8_t*)data, g0 + rows + cols);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
int kf_nf4_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out) {
    KF_REQUIRE(ctx, w->gama_dev && w->cols % 2 == 0, "NF4 tensor needs its gama (codebooks)");
    kf_nf4_dequant_kernel<<<w->rows, 256, 0, ctx->stream>>>((const uint8_t*)w->data_dev, nf4_lut(*w), w->cols, (uint16_t*)out);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
int kf_nf4_embed(kf_ctx* ctx, void* out, const kf_tensor_desc* w, const int32_t* tokens, int M) {
    KF_REQUIRE(ctx, w->gama_dev && w->cols % 2 == 0, "NF4 tensor needs its gama (codebooks)");
    dim3 grid((w->cols / 2 + 255) / 256, M);
    kf_nf4_embed_kernel<<<grid, 256, 0, ctx->stream>>>((uint16_t*)out, (const uint8_t*)w->data_dev, nf4_lut(*w), tokens, w->rows, w->cols);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
// M <= 8
int kf_nf4_gemv(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    KF_REQUIRE(ctx, w->gama_dev && w->cols % 32 == 0 && M >= 1 && M <= 8 && (((uintptr_t)w->data_dev | (uintptr_t)x) & 15) == 0,
               "NF4 GEMV: K a multiple of 32, up to 8 tokens, 16-byte aligned operands");
    KF_REQUIRE(ctx, epilogue == KF_EPI_NONE || epilogue == KF_EPI_F32 || (epilogue == KF_EPI_RESIDUAL && residual), "epilogue");
    const dim3 grid((w->rows + 7) / 8);
#define KF_NF4_GO(MT)                                                                                                                       \
    KF_CUDA(ctx, kf_launch_pdl(ctx, kf_nf4_gemv_kernel<MT>, grid, dim3(256), 0, y, (const uint8_t*)w->data_dev, nf4_lut(*w), (const uint16_t*)x, \
                               (const uint16_t*)residual, w->rows, w->cols, M, epilogue))
    if (M == 1)
        KF_NF4_GO(1);
    else if (M == 2)
        KF_NF4_GO(2);
    else if (M <= 4)
        KF_NF4_GO(4);
    else
        KF_NF4_GO(8);
#undef KF_NF4_GO
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
