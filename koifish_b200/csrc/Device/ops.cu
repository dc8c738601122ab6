// ops.cu -- the small decode ops around the matmuls: RMSNorm, fused QK-norm + RoPE + KV append, SwiGLU, residual add,
// embedding gather, greedy argmax.  Reference kernels: see each function.
#include <math.h>

#include <vector>

#include "kf_common.cuh"

// ---------------------------------------------------------------------------------------------------------------- RMSNorm
// rms_norm_kernel / CU_rms_infer (reference src/Device/CUDA/kernel/layernorm.cuh:801-859): y = (x * rsqrt(mean(x^2)+eps)) * w, fp32
// math, bf16 RN.  The reference runs ONE block for the single decode row; here one block per row (batched decode), 16-byte loads.
__global__ void __launch_bounds__(256) kf_rmsnorm_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ x, const uint16_t* __restrict__ w,
                                                          int dim, float eps) {
    __shared__ float red[32];
    const uint16_t* xr = x + (size_t)blockIdx.x * dim;
    uint16_t* orow     = out + (size_t)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float a = bf16lo(q[j]), b = bf16hi(q[j]);
            ss = fmaf(a, a, ss), ss = fmaf(b, b, ss);
        }
    }
    ss = block_sum(ss, red);
    const float s = 1.0f / sqrtf(fmaf(ss, 1.0f / (float)dim, eps));
    for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
        const uint4 g = *reinterpret_cast<const uint4*>(w + i);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = pack_bf16x2((bf16lo(q[j]) * s) * bf16lo(gw[j]), (bf16hi(q[j]) * s) * bf16hi(gw[j]));
        *reinterpret_cast<uint4*>(orow + i) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
extern "C" int kf_rmsnorm(kf_ctx* ctx, void* out, const void* x, const void* w, int rows, int dim, float eps) {
    if (!ctx || !out || !x || !w) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, rows >= 1 && dim >= 8 && dim % 8 == 0, "dim must be a multiple of 8 (the reference requires even, layernorm.cuh:851)");
    kf_rmsnorm_kernel<<<rows, 256, 0, ctx->stream>>>((uint16_t*)out, (const uint16_t*)x, (const uint16_t*)w, dim, eps);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// ---------------------------------------------------------------------------------------------------------------- RoPE table
// inv_freq = 1/powf(theta, 2j/hd), angle = pos*inv_freq (CU_rope2_v0, reference src/Device/CUDA/kernel/operator.cuh:735-772).  The
// (cos, sin) pairs are tabulated once on the host with libm, so the device rotation is free of fast-math transcendentals.
extern "C" int kf_rope_table(kf_ctx* ctx, void* table_dev, int max_seq, int head_dim, float theta) {
    if (!ctx || !table_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, max_seq > 0 && head_dim > 0 && head_dim % 2 == 0, "shape");
    const int half = head_dim / 2;
    std::vector<float2> t((size_t)max_seq * half);
    for (int j = 0; j < half; j++) {
        const float inv_freq = 1.0f / powf(theta, (float)(j * 2) / (float)head_dim);
        for (int pos = 0; pos < max_seq; pos++) {
            const float angle = (float)pos * inv_freq;
            t[(size_t)pos * half + j] = make_float2(cosf(angle), sinf(angle));
        }
    }
    KF_CUDA(ctx, cudaMemcpyAsync(table_dev, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // t goes out of scope
    return KF_OK;
}

// ---------------------------------------------------------------------------------------------------------------- QK-norm + RoPE + KV
// ROPE::cuInfer (reference src/Device/CUDA/kernel/rope.cu:645-672): per-head RMSNorm of q and k (CU_rmsnorm_multihead semantics,
// layernorm.cuh:750-798: fp32, bf16 RN, eps 1e-6), then CU_rope2_v0 on both; K and V land in row `pos` of the layer's cache (the
// reference lets the K/V projections write there directly, src/Manifold/TGraph.cpp:198-208).  One block per (token, head); thread j
// owns the rotation pair (j, j + hd/2).  Rounding points are those of the reference's separate kernels (norm -> bf16 -> rope -> bf16),
// with RN where the reference rounds stochastically.
__global__ void kf_qknorm_rope_kv_kernel(uint16_t* __restrict__ q, const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                                         const uint16_t* __restrict__ qw, const uint16_t* __restrict__ kw, uint16_t* __restrict__ kcache,
                                         uint16_t* __restrict__ vcache, const float2* __restrict__ table, const int32_t* __restrict__ pos_dev,
                                         int n_head, int n_kv, int hd, float eps, size_t seq_stride) {
    __shared__ float red[32];
    const int m = blockIdx.y, h = blockIdx.x, j = threadIdx.x, half = hd / 2;
    const int pos   = pos_dev[m];
    const bool is_q = h < n_head;
    const int kvh   = h - n_head;
    const uint16_t* src = is_q ? q + ((size_t)m * n_head + h) * hd : k + ((size_t)m * n_kv + kvh) * hd;
    const uint16_t* nw  = is_q ? qw : kw;
    uint16_t* dst       = is_q ? q + ((size_t)m * n_head + h) * hd : kcache + (size_t)m * seq_stride + ((size_t)pos * n_kv + kvh) * hd;
    const float x1 = bf16_bits_to_f32(src[j]), x2 = bf16_bits_to_f32(src[j + half]);
    float ss = fmaf(x1, x1, 0.f);
    ss       = fmaf(x2, x2, ss);
    ss       = block_sum(ss, red);
    float n1 = x1, n2 = x2;
    if (nw) {  // Qwen3 isQKNormal (reference src/Transformer/QWen.cpp:16-58)
        const float s = 1.0f / sqrtf(fmaf(ss, 1.0f / (float)hd, eps));
        n1 = bf16_bits_to_f32(f32_to_bf16_bits((x1 * s) * bf16_bits_to_f32(nw[j])));
        n2 = bf16_bits_to_f32(f32_to_bf16_bits((x2 * s) * bf16_bits_to_f32(nw[j + half])));
    }
    const float2 cs = table[(size_t)pos * half + j];
    dst[j]          = f32_to_bf16_bits(fmaf(n1, cs.x, -(n2 * cs.y)));
    dst[j + half]   = f32_to_bf16_bits(fmaf(n1, cs.y, n2 * cs.x));
    if (!is_q) {  // the V row of this kv head
        const uint16_t* vs = v + ((size_t)m * n_kv + kvh) * hd;
        uint16_t* vd       = vcache + (size_t)m * seq_stride + ((size_t)pos * n_kv + kvh) * hd;
        vd[j] = vs[j], vd[j + half] = vs[j + half];
    }
}
extern "C" int kf_qknorm_rope_kvappend(kf_ctx* ctx, void* q, const void* k, const void* v, const void* qw, const void* kw, void* kcache,
                                       void* vcache, const void* table, const int32_t* pos_dev, int M, int n_head, int n_kv, int hd,
                                       int max_seq, float eps, size_t seq_stride) {
    if (!ctx || !q || !k || !v || !kcache || !vcache || !table || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1 && hd % 2 == 0 && hd / 2 <= 1024 && n_head % n_kv == 0 && max_seq > 0, "shape");
    if (M >= 16 && (hd == 128 || hd == 64))  // big panels: one warp per (token, head), attention.cu
        return kf_qknorm_rope_kv_warp(ctx, q, k, v, qw, kw, kcache, vcache, table, pos_dev, M, n_head, n_kv, hd, eps, seq_stride);
    dim3 grid(n_head + n_kv, M);
    kf_qknorm_rope_kv_kernel<<<grid, hd / 2, 0, ctx->stream>>>((uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v, (const uint16_t*)qw,
                                                              (const uint16_t*)kw, (uint16_t*)kcache, (uint16_t*)vcache, (const float2*)table,
                                                              pos_dev, n_head, n_kv, hd, eps, seq_stride);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// ---------------------------------------------------------------------------------------------------------------- elementwise
// CU_swiglu_v0 (reference src/Device/CUDA/Activation.cu:86-93)
__global__ void __launch_bounds__(256) kf_swiglu_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ gate, const uint16_t* __restrict__ up, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float g = bf16_bits_to_f32(gate[i]), u = bf16_bits_to_f32(up[i]);
        out[i]        = f32_to_bf16_bits((g * u) / (1.0f + expf(-g)));
    }
}
// the same arithmetic, 8 elements (16 bytes) per thread: the prefill path runs it over tokens x ffn elements (HBM-bound)
__global__ void __launch_bounds__(256) kf_swiglu_vec_kernel(uint4* __restrict__ out, const uint4* __restrict__ gate, const uint4* __restrict__ up, size_t n8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 g4 = __ldg(gate + i), u4 = __ldg(up + i);
    const uint32_t g[4] = {g4.x, g4.y, g4.z, g4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float g0 = bf16lo(g[j]), g1 = bf16hi(g[j]), u0 = bf16lo(u[j]), u1 = bf16hi(u[j]);
        o[j] = pack_bf16x2((g0 * u0) / (1.0f + expf(-g0)), (g1 * u1) / (1.0f + expf(-g1)));
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
// CU_add3 (reference src/Device/CUDA/kernel/packedN.cuh:867-875): fp32 add, bf16 out (RN here, stochastic there)
__global__ void __launch_bounds__(256) kf_add_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f32_to_bf16_bits(bf16_bits_to_f32(a[i]) + bf16_bits_to_f32(b[i]));
}
__global__ void __launch_bounds__(256) kf_residual_add_f32_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ res, const float* __restrict__ sum, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f32_to_bf16_bits(bf16_bits_to_f32(res[i]) + bf16_bits_to_f32(f32_to_bf16_bits(sum[i])));
}
// d = alpha * acc + beta * d + bias[row]: the epilogue of TASKA_AxB (src/Tensor/GTensor.hpp:698-741, cuBLASLt computes it in fp32 and
// rounds once to bf16)
__global__ void __launch_bounds__(256) kf_axb_epilogue_kernel(uint16_t* __restrict__ d, const float* __restrict__ acc, const uint16_t* __restrict__ bias,
                                                              float alpha, float beta, int rows, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = alpha * acc[i];
    if (beta != 0.f) v = fmaf(beta, bf16_bits_to_f32(d[i]), v);
    if (bias) v += bf16_bits_to_f32(bias[i % rows]);
    d[i] = f32_to_bf16_bits(v);
}
__global__ void kf_advance_pos_kernel(int32_t* pos, int M) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) pos[i] += 1;
}
extern "C" int kf_residual_add_f32(kf_ctx* ctx, void* out, const void* res, const float* sum, size_t n) {
    if (!ctx || !out || !res || !sum) return KF_ERR_BAD_ARG;
    if (!n) return KF_OK;
    kf_residual_add_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint16_t*)out, (const uint16_t*)res, sum, n);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
// all-gathered vocabulary slices [W][M][vl] -> logits rows [M][W * vl], 16 bytes per thread
__global__ void __launch_bounds__(256) kf_relayout_wmv_kernel(uint4* __restrict__ out, const uint4* __restrict__ in, int W, int M, int vl8) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)W * M * vl8;
    if (i >= n) return;
    const int j = (int)(i % vl8), m = (int)((i / vl8) % M), r = (int)(i / ((size_t)vl8 * M));
    out[((size_t)m * W + r) * vl8 + j] = in[i];
}
extern "C" int kf_relayout_wmv(kf_ctx* ctx, void* out, const void* in, int W, int M, int vl) {
    if (!ctx || !out || !in || out == in) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, vl % 8 == 0 && (((uintptr_t)out | (uintptr_t)in) & 15) == 0, "vocabulary slice must be a multiple of 8 bf16, 16-byte aligned");
    const size_t n = (size_t)W * M * (vl / 8);
    kf_relayout_wmv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint4*)out, (const uint4*)in, W, M, vl / 8);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
int kf_axb_epilogue(kf_ctx* ctx, void* d, const float* acc, const void* bias, float alpha, float beta, int rows, size_t n) {
    kf_axb_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint16_t*)d, acc, (const uint16_t*)bias, alpha, beta, rows, n);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
extern "C" int kf_advance_pos(kf_ctx* ctx, int32_t* pos, int M) {
    if (!ctx || !pos || M < 1) return KF_ERR_BAD_ARG;
    kf_advance_pos_kernel<<<(M + 63) / 64, 64, 0, ctx->stream>>>(pos, M);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
extern "C" int kf_swiglu(kf_ctx* ctx, void* out, const void* gate, const void* up, size_t n) {
    if (!ctx || !out || !gate || !up) return KF_ERR_BAD_ARG;
    if (!n) return KF_OK;
    if (n % 8 == 0 && (((uintptr_t)out | (uintptr_t)gate | (uintptr_t)up) & 15) == 0)
        kf_swiglu_vec_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, ctx->stream>>>((uint4*)out, (const uint4*)gate, (const uint4*)up, n / 8);
    else
        kf_swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint16_t*)out, (const uint16_t*)gate, (const uint16_t*)up, n);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
extern "C" int kf_add(kf_ctx* ctx, void* out, const void* a, const void* b, size_t n) {
    if (!ctx || !out || !a || !b) return KF_ERR_BAD_ARG;
    if (!n) return KF_OK;
    kf_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint16_t*)out, (const uint16_t*)a, (const uint16_t*)b, n);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// ---------------------------------------------------------------------------------------------------------------- embedding
// TokenEmbed::cuInfer (reference src/Device/CUDA/NeuronFuse.cu:176-207; CU_embed_forw_1 embed.cuh:111-120): out[m] = W[token[m]],
// dequantised on the fly when the table is stored in 8 bits or in packed 128-bit words.
__global__ void __launch_bounds__(256) kf_embed_kernel(uint16_t* __restrict__ out, const uint8_t* __restrict__ data, const uint16_t* __restrict__ gZero,
                                                        const uint16_t* __restrict__ gStep, const int32_t* __restrict__ tokens, int rows, int cols,
                                                        int type, int bits, int group, int qbias, int deq_fma) {
    const int m = blockIdx.y;
    int tok     = tokens[m];
    tok         = tok < 0 ? 0 : (tok >= rows ? rows - 1 : tok);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const size_t e = (size_t)tok * cols + c;
    uint16_t r;
    if (type == KF_T_BF16) {
        r = reinterpret_cast<const uint16_t*>(data)[e];
    } else if (type == KF_T_F8E5M2) {
        r = f32_to_bf16_bits(__half2float(__ushort_as_half((unsigned short)((unsigned short)data[e] << 8))));
    } else {
        const int per = 128 / bits, half = per / 2;
        const size_t wi = e / per;
        const int j     = (int)(e % per);
        const unsigned long long* wp = reinterpret_cast<const unsigned long long*>(data + 16 * wi);
        const unsigned long long src = j < half ? wp[1] : wp[0];  // {low, high}
        const int jj   = j < half ? j : j - half;
        const int code = (int)((src >> (64 - bits * (jj + 1))) & ((1u << bits) - 1));
        const size_t g = e / group;
        r = deq_fma ? kf_deq_scalar<true>(code - qbias, gStep[g], gZero[g]) : kf_deq_scalar<false>(code - qbias, gStep[g], gZero[g]);
    }
    out[(size_t)m * cols + c] = r;
}
extern "C" int kf_embed(kf_ctx* ctx, void* out, const kf_tensor_desc* w, const int32_t* tokens, int M) {
    if (!ctx || !out || !w || !tokens || !w->data_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1, "M");
    if (w->type == KF_T_NF4) return kf_nf4_embed(ctx, out, w, tokens, M);
    KF_REQUIRE(ctx, w->type != KF_T_AWQ4, "embedding tables in the AWQ layout are not supported");
    const int bits = kf_type_bits(w->type);
    KF_REQUIRE(ctx, bits > 0, "type");
    const uint16_t *gz = nullptr, *gs = nullptr;
    if (kf_type_packed(w->type)) {
        KF_REQUIRE(ctx, kf_has_gama(*w) && w->group > 0, "packed embedding needs gama");
        gz = kf_gama_zero(*w), gs = kf_gama_step(*w);
    }
    dim3 grid((w->cols + 255) / 256, M);
    kf_embed_kernel<<<grid, 256, 0, ctx->stream>>>((uint16_t*)out, (const uint8_t*)w->data_dev, gz, gs, tokens, w->rows, w->cols, w->type, bits,
                                                   w->group, w->qbias, ctx->deq_fma);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// ---------------------------------------------------------------------------------------------------------------- argmax
// Greedy sampling on the device (the reference copies the logits to the host and samples there, src/Manifold/GoPT.cpp:614-630).
// Ties resolve to the lowest index (numpy / torch argmax convention).
__global__ void __launch_bounds__(1024) kf_argmax_kernel(int32_t* __restrict__ out, const uint16_t* __restrict__ logits, int vocab) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const uint16_t* row = logits + (size_t)blockIdx.x * vocab;
    float best = -INFINITY;
    int bi     = 0x7fffffff;
    // 16-byte loads, four per thread in flight (152K logits: ~5 us instead of ~70 with 2-byte loads); a thread visits its indices in
    // increasing order, so "first maximum wins" needs only the strict comparison here
    const int nvec = (vocab % 8 == 0 && ((uintptr_t)row & 15) == 0) ? vocab / 8 : 0;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += blockDim.x * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x;
            v[u]        = i < nvec ? __ldg(reinterpret_cast<const uint4*>(row) + i) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int base      = (i0 + u * blockDim.x) * 8;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float a = bf16lo(w[j]), b = bf16hi(w[j]);
                if (a > best) best = a, bi = base + 2 * j;
                if (b > best) best = b, bi = base + 2 * j + 1;
            }
        }
    }
    for (int i = nvec * 8 + threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = bf16_bits_to_f32(row[i]);
        if (v > best || (v == best && i < bi)) best = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi   = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sv[warp] = best, si[warp] = bi;
    __syncthreads();
    if (warp == 0) {
        best = lane < (int)(blockDim.x >> 5) ? sv[lane] : -INFINITY;
        bi   = lane < (int)(blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi   = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
        }
        if (lane == 0) out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
    }
}
extern "C" int kf_argmax(kf_ctx* ctx, int32_t* out, const void* logits, int M, int vocab) {
    if (!ctx || !out || !logits) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1 && vocab >= 1, "shape");
    kf_argmax_kernel<<<M, 1024, 0, ctx->stream>>>(out, (const uint16_t*)logits, vocab);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
