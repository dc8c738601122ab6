// attention.cu -- GQA decode attention over a contiguous bf16 KV cache, split-K over the sequence (flash-decoding).
//
// Replaces the reference's three launches  attention_qk_kernel<<<nHead, min(1024,pos+1)>>>  +  CU_softmax_multihead<<<nHead,1>>>
// +  attention_v_kernel<<<nHead,hd>>>  (src/Device/CUDA/kernel/operator.cuh:573-632, 252-277, 650-668; called from
// SelfAttention::cuInfer, src/Device/CUDA/QKV.cu:668-674), which silently drop positions >= 1024 and run the softmax on one
// thread.  Math follows the fp32-score variant (pipe path, src/Device/CUDA/Generate.cu:273): s = (q.k)/sqrt(hd) in fp32,
// softmax in fp32, out = sum p*v in fp32, bf16 RN.  kv head = h / (n_head / n_kv).
//
// grid (n_head, M, nsplit), 4 warps per CTA.  A warp walks the tokens of its slice; each lane owns hd/32 contiguous dims, so one
// K (or V) row of a head is one coalesced 128/256-byte warp load.  Online softmax per warp; the 4 warps and then the splits are
// merged with the usual (max, sum, acc) rescaling, in fixed order.
#include <algorithm>

#include "kf_common.cuh"

namespace {
constexpr int kAttnWarps = 4;

template <int DPL>  // dims per lane: 4 (hd 128) or 2 (hd 64)
__device__ __forceinline__ void load_row(float (&f)[DPL], const uint16_t* p) {
    if constexpr (DPL == 4) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        f[0] = bf16lo(v.x), f[1] = bf16hi(v.x), f[2] = bf16lo(v.y), f[3] = bf16hi(v.y);
    } else {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
        f[0] = bf16lo(v), f[1] = bf16hi(v);
    }
}

// partial layout in the workspace: [M][n_head][nsplit][hd + 2] floats : acc[hd], max, sum
template <int DPL>
__global__ void __launch_bounds__(kAttnWarps * 32) kf_attn_decode_kernel(uint16_t* __restrict__ out, float* __restrict__ ws,
                                                                         const uint16_t* __restrict__ q, const uint16_t* __restrict__ kc,
                                                                         const uint16_t* __restrict__ vc, const int32_t* __restrict__ pos_dev,
                                                                         int n_head, int n_kv, int nsplit, float sqrt_hd, size_t seq_stride) {
    constexpr int HD = DPL * 32;
    __shared__ float s_acc[kAttnWarps][HD];
    __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
    const int h = blockIdx.x, m = blockIdx.y, split = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kvh = h / (n_head / n_kv), kv_dim = n_kv * HD;
    const int len = pos_dev[m] + 1;
    const int t0 = (int)(((long long)split * len) / nsplit), t1 = (int)(((long long)(split + 1) * len) / nsplit);

    float qf[DPL];
    load_row<DPL>(qf, q + ((size_t)m * n_head + h) * HD + lane * DPL);
    float mx = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0.f;

    const uint16_t* kbase = kc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    const uint16_t* vbase = vc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    for (int t = t0 + warp; t < t1; t += kAttnWarps) {
        float kf[DPL], vf[DPL];
        load_row<DPL>(kf, kbase + (size_t)t * kv_dim);
        load_row<DPL>(vf, vbase + (size_t)t * kv_dim);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DPL; d++) s = fmaf(qf[d], kf[d], s);
        s = warp_sum(s) / sqrt_hd;  // the reference divides (operator.cuh:630)
        const float mn = fmaxf(mx, s);
        const float c  = expf(mx - mn);  // 0 on the first token (mx = -inf)
        const float p  = expf(s - mn);
        l = l * c + p;
#pragma unroll
        for (int d = 0; d < DPL; d++) acc[d] = fmaf(p, vf[d], acc[d] * c);
        mx = mn;
    }
    // merge the warps of this CTA (fixed order)
#pragma unroll
    for (int d = 0; d < DPL; d++) s_acc[warp][lane * DPL + d] = acc[d];
    if (lane == 0) s_m[warp] = mx, s_l[warp] = l;
    __syncthreads();
    if (warp == 0) {
        float M_ = -INFINITY;
#pragma unroll
        for (int w = 0; w < kAttnWarps; w++) M_ = fmaxf(M_, s_m[w]);
        float L_ = 0.f, o[DPL];
#pragma unroll
        for (int d = 0; d < DPL; d++) o[d] = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnWarps; w++) {
            const float c = s_m[w] == -INFINITY ? 0.f : expf(s_m[w] - M_);
            L_ += s_l[w] * c;
#pragma unroll
            for (int d = 0; d < DPL; d++) o[d] = fmaf(s_acc[w][lane * DPL + d], c, o[d]);
        }
        if (nsplit == 1) {
            const float inv = 1.0f / L_;
            uint16_t* op    = out + ((size_t)m * n_head + h) * HD + lane * DPL;
#pragma unroll
            for (int d = 0; d < DPL; d++) op[d] = f32_to_bf16_bits(o[d] * inv);
        } else {
            float* wp = ws + (((size_t)m * n_head + h) * nsplit + split) * (HD + 2);
#pragma unroll
            for (int d = 0; d < DPL; d++) wp[lane * DPL + d] = o[d];
            if (lane == 0) wp[HD] = M_, wp[HD + 1] = L_;
        }
    }
}

template <int DPL>
__global__ void kf_attn_combine_kernel(uint16_t* __restrict__ out, const float* __restrict__ ws, int n_head, int nsplit) {
    constexpr int HD = DPL * 32;
    const int h = blockIdx.x, m = blockIdx.y, d = threadIdx.x;
    const float* wp = ws + ((size_t)m * n_head + h) * nsplit * (HD + 2);
    float M_ = -INFINITY;
    for (int s = 0; s < nsplit; s++) M_ = fmaxf(M_, wp[s * (HD + 2) + HD]);
    float L_ = 0.f, o = 0.f;
    for (int s = 0; s < nsplit; s++) {
        const float ms = wp[s * (HD + 2) + HD];
        const float c  = ms == -INFINITY ? 0.f : expf(ms - M_);
        L_ += wp[s * (HD + 2) + HD + 1] * c;
        o = fmaf(wp[s * (HD + 2) + d], c, o);
    }
    out[((size_t)m * n_head + h) * HD + d] = f32_to_bf16_bits(o * (1.0f / L_));
}
}  // namespace

extern "C" int kf_attn_decode(kf_ctx* ctx, void* out, const void* q, const void* kc, const void* vc, const int32_t* pos_dev, int M, int n_head,
                              int n_kv, int hd, int max_seq, int max_pos_hint, size_t seq_stride) {
    if (!ctx || !out || !q || !kc || !vc || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (hd == 128 || hd == 64) && n_head % n_kv == 0 && M >= 1 && max_seq >= 1, "head_dim 64/128, GQA");
    // enough CTAs to cover the SMs, at least ~32 tokens per warp-slice
    int nsplit = ctx->attn_split;
    if (nsplit <= 0) {
        const int len = std::max(1, std::min(max_seq, max_pos_hint + 1));
        nsplit        = (2 * ctx->sm_count + n_head * M - 1) / (n_head * M);
        nsplit        = std::min(nsplit, std::max(1, len / (kAttnWarps * 16)));
        nsplit        = std::max(1, std::min(nsplit, 64));
    }
    float* ws = nullptr;
    if (nsplit > 1) {
        int rc = kf_ensure_attn_ws(ctx, (size_t)M * n_head * nsplit * (hd + 2) * sizeof(float));
        if (rc) return rc;
        ws = ctx->attn_ws;
    }
    dim3 grid(n_head, M, nsplit);
    const float isq = sqrtf((float)hd);
    if (hd == 128)
        kf_attn_decode_kernel<4><<<grid, kAttnWarps * 32, 0, ctx->stream>>>((uint16_t*)out, ws, (const uint16_t*)q, (const uint16_t*)kc,
                                                                            (const uint16_t*)vc, pos_dev, n_head, n_kv, nsplit, isq, seq_stride);
    else
        kf_attn_decode_kernel<2><<<grid, kAttnWarps * 32, 0, ctx->stream>>>((uint16_t*)out, ws, (const uint16_t*)q, (const uint16_t*)kc,
                                                                            (const uint16_t*)vc, pos_dev, n_head, n_kv, nsplit, isq, seq_stride);
    KF_LAUNCH_CHECK(ctx);
    if (nsplit > 1) {
        dim3 g2(n_head, M);
        if (hd == 128)
            kf_attn_combine_kernel<4><<<g2, 128, 0, ctx->stream>>>((uint16_t*)out, ws, n_head, nsplit);
        else
            kf_attn_combine_kernel<2><<<g2, 64, 0, ctx->stream>>>((uint16_t*)out, ws, n_head, nsplit);
        KF_LAUNCH_CHECK(ctx);
    }
    return KF_OK;
}
