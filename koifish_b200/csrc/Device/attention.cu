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
constexpr int kTokBatch  = 8;  // cache rows in flight per warp (fused kernel: the first batch is requested before the dependency wait)

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
    // kTokBatch rows of K and V are requested before any of them is used: the loop is latency bound otherwise
    for (int tb = t0 + warp * kTokBatch; tb < t1; tb += kAttnWarps * kTokBatch) {
        float kf[kTokBatch][DPL], vf[kTokBatch][DPL];
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            const int t = min(tb + u, t1 - 1);  // clamped rows are loaded but not used
            load_row<DPL>(kf[u], kbase + (size_t)t * kv_dim);
            load_row<DPL>(vf[u], vbase + (size_t)t * kv_dim);
        }
        float s[kTokBatch];
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            s[u] = 0.f;
#pragma unroll
            for (int d = 0; d < DPL; d++) s[u] = fmaf(qf[d], kf[u][d], s[u]);
        }
#pragma unroll
        for (int o_ = 16; o_ > 0; o_ >>= 1)
#pragma unroll
            for (int u = 0; u < kTokBatch; u++) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o_);
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            if (tb + u >= t1) break;
            const float sc = s[u] / sqrt_hd;  // the reference divides (operator.cuh:630)
            const float mn = fmaxf(mx, sc);
            const float c  = expf(mx - mn);  // 0 on the first token (mx = -inf)
            const float p  = expf(sc - mn);
            l = l * c + p;
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = fmaf(p, vf[u][d], acc[d] * c);
            mx = mn;
        }
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
// ---- one launch for ROPE::cuInfer + the three attention kernels (decode: every token of the launch belongs to its own sequence) ----
// Each CTA (head h, token m, split) re-derives the normalised + rotated q of its head and the k / v of the current position in
// registers (128 elements: cheaper than a launch), attends to the cached positions < pos of its slice plus -- in the last slice --
// the current position straight from registers, and the last CTA to arrive for (m, h) merges the slices in fixed order.  The K / V
// rows of position pos are written to the cache by one designated CTA per kv head; nobody reads row pos from the cache here.
template <int DPL>
__device__ __forceinline__ void norm_rope_row(float (&o)[DPL], const uint16_t* src, const uint16_t* nw, const float2* cs_row, int lane, float eps) {
    constexpr int HD = DPL * 32;
    float x[DPL];
    load_row<DPL>(x, src + lane * DPL);
    if (nw) {
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < DPL; d++) ss = fmaf(x[d], x[d], ss);
        ss            = warp_sum(ss);
        const float s = 1.0f / sqrtf(fmaf(ss, 1.0f / (float)HD, eps));
        float w[DPL];
        load_row<DPL>(w, nw + lane * DPL);
#pragma unroll
        for (int d = 0; d < DPL; d++) x[d] = bf16_bits_to_f32(f32_to_bf16_bits((x[d] * s) * w[d]));  // the norm kernel's bf16 output
    }
    // half-split rotation: dim j pairs with j + HD/2, held by lane ^ 16
#pragma unroll
    for (int d = 0; d < DPL; d++) {
        const float other = __shfl_xor_sync(0xffffffffu, x[d], 16);
        const int j       = (lane & 15) * DPL + d;
        const float2 cs   = cs_row[j];
        const float r     = lane < 16 ? fmaf(x[d], cs.x, -(other * cs.y)) : fmaf(other, cs.y, x[d] * cs.x);
        o[d]              = bf16_bits_to_f32(f32_to_bf16_bits(r));
    }
}

template <int DPL>
__global__ void __launch_bounds__(kAttnWarps * 32) kf_attn_fused_kernel(uint16_t* __restrict__ out, float* __restrict__ ws, unsigned* __restrict__ cnt,
                                                                        const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                                        const uint16_t* __restrict__ v, const uint16_t* __restrict__ qw,
                                                                        const uint16_t* __restrict__ kw, uint16_t* __restrict__ kc,
                                                                        uint16_t* __restrict__ vc, const float2* __restrict__ table,
                                                                        const int32_t* __restrict__ pos_dev, int n_head, int n_kv, int nsplit,
                                                                        float sqrt_hd, float eps, size_t seq_stride) {
    constexpr int HD = DPL * 32;
    __shared__ float s_acc[kAttnWarps][HD];
    __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
    __shared__ int s_last;
    const int h = blockIdx.x, m = blockIdx.y, split = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = n_head / n_kv, kvh = h / group, kv_dim = n_kv * HD;
    kf_grid_launch_dependents();
    // Everything up to the wait is independent of the QKV GEMV that precedes this kernel: the position, the slice and the FIRST batch
    // of cached K / V rows (rows < pos were written by earlier tokens).  Under programmatic dependent launch these loads overlap the
    // predecessor's tail.
    const int pos = pos_dev[m], len = pos + 1;
    const int t0 = (int)(((long long)split * len) / nsplit), t1 = (int)(((long long)(split + 1) * len) / nsplit);
    const float2* cs_row  = table + (size_t)pos * (HD / 2);
    const uint16_t* kbase = kc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    const uint16_t* vbase = vc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    float kf[kTokBatch][DPL], vf[kTokBatch][DPL];
    auto load_batch = [&](int tb_) {
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            const int t = tb_ + u;
            if (t < t1 && t != pos) {
                load_row<DPL>(kf[u], kbase + (size_t)t * kv_dim);
                load_row<DPL>(vf[u], vbase + (size_t)t * kv_dim);
            }
        }
    };
    int tb = t0 + warp * kTokBatch;
    if (tb < t1) load_batch(tb);
    kf_grid_dependency_wait();  // q / k / v come from the QKV GEMV right before us

    float qf[DPL], knew[DPL], vnew[DPL];
    norm_rope_row<DPL>(qf, q + ((size_t)m * n_head + h) * HD, qw, cs_row, lane, eps);
    const bool has_new = t1 == len;  // the last slice owns the current position
    if (has_new) {
        norm_rope_row<DPL>(knew, k + ((size_t)m * n_kv + kvh) * HD, kw, cs_row, lane, eps);
        load_row<DPL>(vnew, v + ((size_t)m * n_kv + kvh) * HD + lane * DPL);
        if (h % group == 0 && warp == 0) {  // append K / V of this kv head at row pos (KVCache, Cache.cpp:43-58)
            uint16_t* kd = kc + (size_t)m * seq_stride + (size_t)pos * kv_dim + (size_t)kvh * HD + lane * DPL;
            uint16_t* vd = vc + (size_t)m * seq_stride + (size_t)pos * kv_dim + (size_t)kvh * HD + lane * DPL;
#pragma unroll
            for (int d = 0; d < DPL; d++) kd[d] = f32_to_bf16_bits(knew[d]), vd[d] = f32_to_bf16_bits(vnew[d]);
        }
    }
    float mx = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0.f;
    for (; tb < t1; tb += kAttnWarps * kTokBatch) {
#pragma unroll
        for (int u = 0; u < kTokBatch; u++)
            if (tb + u == pos) {  // the current position comes from registers, never from the cache
#pragma unroll
                for (int d = 0; d < DPL; d++) kf[u][d] = knew[d], vf[u][d] = vnew[d];
            }
        float s[kTokBatch];
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            s[u] = 0.f;
#pragma unroll
            for (int d = 0; d < DPL; d++) s[u] = fmaf(qf[d], kf[u][d], s[u]);
        }
#pragma unroll
        for (int o_ = 16; o_ > 0; o_ >>= 1)
#pragma unroll
            for (int u = 0; u < kTokBatch; u++) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o_);
#pragma unroll
        for (int u = 0; u < kTokBatch; u++) {
            if (tb + u >= t1) break;
            const float sc = s[u] / sqrt_hd;
            const float mn = fmaxf(mx, sc);
            const float c  = expf(mx - mn);
            const float p  = expf(sc - mn);
            l = l * c + p;
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = fmaf(p, vf[u][d], acc[d] * c);
            mx = mn;
        }
        if (tb + kAttnWarps * kTokBatch < t1) load_batch(tb + kAttnWarps * kTokBatch);
    }
#pragma unroll
    for (int d = 0; d < DPL; d++) s_acc[warp][lane * DPL + d] = acc[d];
    if (lane == 0) s_m[warp] = mx, s_l[warp] = l;
    __syncthreads();
    float M_ = -INFINITY, L_ = 0.f, o[DPL];
    if (warp == 0) {
#pragma unroll
        for (int w = 0; w < kAttnWarps; w++) M_ = fmaxf(M_, s_m[w]);
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
            for (int d = 0; d < DPL; d++) __stcg(wp + lane * DPL + d, o[d]);
            if (lane == 0) __stcg(wp + HD, M_), __stcg(wp + HD + 1, L_);
        }
    }
    if (nsplit == 1) return;
    // last CTA of (m, h) merges the slices in split order (deterministic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(cnt + (size_t)m * n_head + h, 1u);
        s_last              = prev == (unsigned)(nsplit - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) {
        const float* wp = ws + ((size_t)m * n_head + h) * nsplit * (HD + 2);
        float Mx = -INFINITY;
        for (int s = 0; s < nsplit; s++) Mx = fmaxf(Mx, __ldcg(wp + s * (HD + 2) + HD));
        float Ls = 0.f, oo[DPL];
#pragma unroll
        for (int d = 0; d < DPL; d++) oo[d] = 0.f;
        for (int s = 0; s < nsplit; s++) {
            const float ms = __ldcg(wp + s * (HD + 2) + HD);
            const float c  = ms == -INFINITY ? 0.f : expf(ms - Mx);
            Ls += __ldcg(wp + s * (HD + 2) + HD + 1) * c;
#pragma unroll
            for (int d = 0; d < DPL; d++) oo[d] = fmaf(__ldcg(wp + s * (HD + 2) + lane * DPL + d), c, oo[d]);
        }
        const float inv = 1.0f / Ls;
        uint16_t* op    = out + ((size_t)m * n_head + h) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; d++) op[d] = f32_to_bf16_bits(oo[d] * inv);
        if (lane == 0) cnt[(size_t)m * n_head + h] = 0u;  // self-reset
    }
}

// ---- the same fused step for contexts up to ~1K tokens: the slices of one (token, head) form a THREAD-BLOCK CLUSTER and are merged
// through distributed shared memory instead of a global workspace + arrival counter + last-CTA pass.  What is left on the critical
// path after the QKV GEMV finishes is one L2 round trip (q / k / v of the new token), the score / softmax arithmetic of 16 cached rows
// per warp -- all of which were requested BEFORE the dependency wait, as packed bf16 -- and one cluster barrier.
template <int DPL>
__device__ __forceinline__ uint2 load_raw(const uint16_t* p) {
    if constexpr (DPL == 4) return *reinterpret_cast<const uint2*>(p);
    return make_uint2(*reinterpret_cast<const uint32_t*>(p), 0u);
}
template <int DPL>
__device__ __forceinline__ void unpack_raw(float (&f)[DPL], uint2 v) {
    f[0] = bf16lo(v.x), f[1] = bf16hi(v.x);
    if constexpr (DPL == 4) f[2] = bf16lo(v.y), f[3] = bf16hi(v.y);
}
#ifndef KF_ATTN_WARPTOK
#define KF_ATTN_WARPTOK 8
#endif
constexpr int kWarpTok    = KF_ATTN_WARPTOK;  // cached rows per warp per pass, all in flight at once, held as packed bf16
constexpr int kMaxCluster = 8;   // portable cluster size
constexpr int kClusterWarpsMax = 8;

template <int DPL>
__global__ void __launch_bounds__(kClusterWarpsMax * 32) kf_attn_cluster_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ q,
                                                                          const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                                                                          const uint16_t* __restrict__ qw, const uint16_t* __restrict__ kw,
                                                                          uint16_t* __restrict__ kc, uint16_t* __restrict__ vc,
                                                                          const float2* __restrict__ table, const int32_t* __restrict__ pos_dev,
                                                                          int n_head, int n_kv, int nsplit, float sqrt_hd, float eps,
                                                                          size_t seq_stride) {
    constexpr int HD = DPL * 32;
    __shared__ float s_acc[kClusterWarpsMax][HD];
    __shared__ float s_m[kClusterWarpsMax], s_l[kClusterWarpsMax];
    const int nwarps = blockDim.x >> 5;
    __shared__ float s_part[kMaxCluster][HD + 2];  // rank 0's copy receives the partial (acc[hd], max, sum) of every slice
    const int h = blockIdx.x, m = blockIdx.y, split = blockIdx.z;  // cluster = (1, 1, nsplit): rank in cluster == split
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = n_head / n_kv, kvh = h / group, kv_dim = n_kv * HD;
    kf_grid_launch_dependents();
    if (nsplit > 1) cluster_arrive();  // matched by the wait in front of the first remote store: by then every CTA of the cluster runs
    // ---- independent of the QKV GEMV: position, slice, rotation table row and ALL cached rows of this warp's first pass ----
    const int pos = pos_dev[m], len = pos + 1;
    const int t0 = (int)(((long long)split * len) / nsplit), t1 = (int)(((long long)(split + 1) * len) / nsplit);
    const float2* cs_row  = table + (size_t)pos * (HD / 2);
    const uint16_t* kbase = kc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    const uint16_t* vbase = vc + (size_t)m * seq_stride + (size_t)kvh * HD + lane * DPL;
    uint2 kr[kWarpTok], vr[kWarpTok];
    auto load_pass = [&](int tb_) {
#pragma unroll
        for (int u = 0; u < kWarpTok; u++) {
            const int t = tb_ + u;
            if (t < t1 && t != pos) kr[u] = load_raw<DPL>(kbase + (size_t)t * kv_dim), vr[u] = load_raw<DPL>(vbase + (size_t)t * kv_dim);
        }
    };
    int tb = t0 + warp * kWarpTok;
    if (tb < t1) load_pass(tb);
    kf_grid_dependency_wait();  // q / k / v come from the QKV GEMV right before us

    float qf[DPL], knew[DPL], vnew[DPL];
    norm_rope_row<DPL>(qf, q + ((size_t)m * n_head + h) * HD, qw, cs_row, lane, eps);
    const bool has_new = t1 == len;  // the last slice owns the current position
    if (has_new) {
        norm_rope_row<DPL>(knew, k + ((size_t)m * n_kv + kvh) * HD, kw, cs_row, lane, eps);
        load_row<DPL>(vnew, v + ((size_t)m * n_kv + kvh) * HD + lane * DPL);
        if (h % group == 0 && warp == 0) {  // append K / V of this kv head at row pos (KVCache, Cache.cpp:43-58)
            uint16_t* kd = kc + (size_t)m * seq_stride + (size_t)pos * kv_dim + (size_t)kvh * HD + lane * DPL;
            uint16_t* vd = vc + (size_t)m * seq_stride + (size_t)pos * kv_dim + (size_t)kvh * HD + lane * DPL;
#pragma unroll
            for (int d = 0; d < DPL; d++) kd[d] = f32_to_bf16_bits(knew[d]), vd[d] = f32_to_bf16_bits(vnew[d]);
        }
    }
    float mx = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0.f;
    for (; tb < t1; tb += nwarps * kWarpTok) {
        // scores of the pass: per-lane partial dot products, then one butterfly over the 16 values
        float s[kWarpTok];
#pragma unroll
        for (int u = 0; u < kWarpTok; u++) {
            float kf[DPL];
            unpack_raw<DPL>(kf, kr[u]);
            if (tb + u == pos) {  // the current position comes from registers, never from the cache
#pragma unroll
                for (int d = 0; d < DPL; d++) kf[d] = knew[d];
            }
            s[u] = 0.f;
#pragma unroll
            for (int d = 0; d < DPL; d++) s[u] = fmaf(qf[d], kf[d], s[u]);
        }
#pragma unroll
        for (int o_ = 16; o_ > 0; o_ >>= 1)
#pragma unroll
            for (int u = 0; u < kWarpTok; u++) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o_);
        // softmax of the pass with ONE rescale (no serial max / exp chain over the tokens)
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u < kWarpTok; u++) {
            s[u] = tb + u < t1 ? s[u] / sqrt_hd : -INFINITY;  // the reference divides (operator.cuh:630)
            mb   = fmaxf(mb, s[u]);
        }
        const float mn = fmaxf(mx, mb);
        const float c  = expf(mx - mn);  // 0 on the first pass (mx = -inf)
        float lp = 0.f, ap[DPL];
#pragma unroll
        for (int d = 0; d < DPL; d++) ap[d] = 0.f;
#pragma unroll
        for (int u = 0; u < kWarpTok; u++) {
            const float p = expf(s[u] - mn);  // exp(-inf) = 0 for the masked tail
            float vf[DPL];
            unpack_raw<DPL>(vf, vr[u]);
            if (tb + u == pos) {
#pragma unroll
                for (int d = 0; d < DPL; d++) vf[d] = vnew[d];
            }
            lp += p;
#pragma unroll
            for (int d = 0; d < DPL; d++) ap[d] = fmaf(p, tb + u < t1 ? vf[d] : 0.f, ap[d]);
        }
        l = l * c + lp;
#pragma unroll
        for (int d = 0; d < DPL; d++) acc[d] = fmaf(acc[d], c, ap[d]);
        mx = mn;
        if (tb + nwarps * kWarpTok < t1) load_pass(tb + nwarps * kWarpTok);
    }
    // ---- merge the warps of this CTA (fixed order) ----
#pragma unroll
    for (int d = 0; d < DPL; d++) s_acc[warp][lane * DPL + d] = acc[d];
    if (lane == 0) s_m[warp] = mx, s_l[warp] = l;
    __syncthreads();
    if (nsplit > 1) cluster_wait();  // every CTA of the cluster has started: its shared memory may be written
    if (warp == 0) {
        float M_ = -INFINITY, L_ = 0.f, o[DPL];
        for (int w = 0; w < nwarps; w++) M_ = fmaxf(M_, s_m[w]);
#pragma unroll
        for (int d = 0; d < DPL; d++) o[d] = 0.f;
        for (int w = 0; w < nwarps; w++) {
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
        } else {  // hand the slice's partial to rank 0 through distributed shared memory
#pragma unroll
            for (int d = 0; d < DPL; d++) st_cluster_f32(&s_part[split][lane * DPL + d], 0, o[d]);
            if (lane == 0) st_cluster_f32(&s_part[split][HD], 0, M_), st_cluster_f32(&s_part[split][HD + 1], 0, L_);
        }
    }
    if (nsplit == 1) return;
    cluster_arrive();  // release: the remote stores above are visible to rank 0 after its wait
    cluster_wait();
    if (split != 0 || warp != 0) return;
    {  // rank 0 merges the slices in split order (deterministic)
        float Mx = -INFINITY;
        for (int sp = 0; sp < nsplit; sp++) Mx = fmaxf(Mx, s_part[sp][HD]);
        float Ls = 0.f, oo[DPL];
#pragma unroll
        for (int d = 0; d < DPL; d++) oo[d] = 0.f;
        for (int sp = 0; sp < nsplit; sp++) {
            const float ms = s_part[sp][HD];
            const float c  = ms == -INFINITY ? 0.f : expf(ms - Mx);
            Ls += s_part[sp][HD + 1] * c;
#pragma unroll
            for (int d = 0; d < DPL; d++) oo[d] = fmaf(s_part[sp][lane * DPL + d], c, oo[d]);
        }
        const float inv = 1.0f / Ls;
        uint16_t* op    = out + ((size_t)m * n_head + h) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; d++) op[d] = f32_to_bf16_bits(oo[d] * inv);
    }
}

template <int DPL>
cudaError_t launch_attn_cluster(kf_ctx* ctx, dim3 grid, int nsplit, int warps, uint16_t* out, const uint16_t* q, const uint16_t* k, const uint16_t* v,
                                const uint16_t* qw, const uint16_t* kw, uint16_t* kc, uint16_t* vc, const float2* table, const int32_t* pos_dev,
                                int n_head, int n_kv, float sqrt_hd, float eps, size_t seq_stride) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = dim3(warps * 32), cfg.dynamicSmemBytes = 0, cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (nsplit > 1) {
        attr[na].id               = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = (unsigned)nsplit;
        na++;
    }
    if (ctx->pdl) {
        attr[na].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr, cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kf_attn_cluster_kernel<DPL>, out, q, k, v, qw, kw, kc, vc, table, pos_dev, n_head, n_kv, nsplit, sqrt_hd, eps,
                              seq_stride);
}

// ---- prefill: causal attention of a panel of M consecutive tokens of ONE sequence over the cache rows [0, pos0 + M) ------------------
// Replaces running the three decode kernels once per prompt token (the reference's Generate loop feeds the prompt token by token,
// src/Manifold/GoPT.cpp:1111-1235).  Flash-attention forward on the legacy tensor-core path (mma.sync m16n8k16 bf16, fp32 accumulate):
// a CTA owns 64 query rows of one head (4 warps x 16 rows, Q fragments in registers), streams 64-token K / V tiles of the head's kv
// group through a cp.async double buffer (16-byte chunks XOR-swizzled by row so every ldmatrix is conflict-free), keeps the running
// (max, sum) per row and the 16 x hd output tile in registers.  Scores s = (q.k)/sqrt(hd) and the softmax are fp32; P is rounded to
// bf16 for the P.V product (the reference's neuron path stores the scores as bf16 as well, TGraph.cpp:123-124).
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
constexpr int kPfQ = 64, kPfKV = 64;
#ifndef KF_PF_BKV
#define KF_PF_BKV 64
#endif
#ifndef KF_PF_MINB
#define KF_PF_MINB 2
#endif

template <int HD, int BKV>
__global__ void __launch_bounds__(128, KF_PF_MINB) kf_attn_prefill_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ q,
                                                              const uint16_t* __restrict__ kc, const uint16_t* __restrict__ vc,
                                                              const int32_t* __restrict__ pos_dev, int M, int n_head, int n_kv, int max_seq,
                                                              float sqrt_hd) {
    constexpr int CH = HD / 8;  // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t pf_smem[];
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(pf_smem);
    const uint32_t sK = sQ + kPfQ * HD * 2;            // two K tiles
    const uint32_t sV = sK + 2 * BKV * HD * 2;       // two V tiles
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;  // heavy (late) query tiles first
    const int h = blockIdx.y, kvh = h / (n_head / n_kv), kv_dim = n_kv * HD, q_dim = n_head * HD;
    const int pos0 = pos_dev[0];
    const int q0 = qt * kPfQ;
    const int kv_len = pos0 + min(M, q0 + kPfQ);  // rows visible to the last query row of this tile
    const int ntiles = (kv_len + BKV - 1) / BKV;
    auto sw = [](int row, int c) { return (uint32_t)((row * CH + (c ^ (row & 7))) * 16); };

    // ---- Q tile + first K / V tile ----
    for (int i = tid; i < kPfQ * CH; i += 128) {
        const int r = i / CH, c = i % CH;
        const bool ok = q0 + r < M;
        cp16(sQ + sw(r, c), q + (size_t)(ok ? q0 + r : 0) * q_dim + (size_t)h * HD + c * 8, ok);
    }
    auto load_kv = [&](int kt, int buf) {
        for (int i = tid; i < BKV * CH; i += 128) {
            const int r = i / CH, c = i % CH;
            const int t = min(kt * BKV + r, max_seq - 1);  // rows past kv_len are masked below; keep the address inside the cache
            cp16(sK + buf * (BKV * HD * 2) + sw(r, c), kc + (size_t)t * kv_dim + (size_t)kvh * HD + c * 8, true);
            cp16(sV + buf * (BKV * HD * 2) + sw(r, c), vc + (size_t)t * kv_dim + (size_t)kvh * HD + c * 8, true);
        }
    };
    load_kv(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    uint32_t qa[HD / 16][4];
    float o[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    const int p_lo = pos0 + q0 + warp * 16 + g, p_hi = p_lo + 8;  // positions of this thread's two query rows
    const float SC = 1.4426950408889634f / sqrt_hd;  // softmax((q.k)/sqrt(hd)) = 2^((q.k - max) * log2(e)/sqrt(hd)) / sum

    for (int kt = 0; kt < ntiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) load_kv(kt + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int kc_ = 0; kc_ < HD / 16; kc_++) ldsm_x4(qa[kc_], sQ + sw(warp * 16 + (lane & 15), kc_ * 2 + (lane >> 4)));
        }
        const uint32_t kb = sK + buf * (BKV * HD * 2), vb = sV + buf * (BKV * HD * 2);
        // ---- S = Q K^T (16 x 64 per warp) ----
        float sc[BKV / 8][4];
#pragma unroll
        for (int j = 0; j < BKV / 8; j++) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
        for (int kc_ = 0; kc_ < HD / 16; kc_++) {
#pragma unroll
            for (int jp = 0; jp < BKV / 16; jp++) {
                uint32_t b[4];
                ldsm_x4(b, kb + sw(jp * 16 + (lane & 7) + (lane >> 4) * 8, kc_ * 2 + ((lane >> 3) & 1)));
                mma_bf16(sc[2 * jp], qa[kc_], b[0], b[1]);
                mma_bf16(sc[2 * jp + 1], qa[kc_], b[2], b[3]);
            }
        }
        // ---- scale, causal mask, online softmax (rows g and g + 8 of the warp's 16) ----
        float mx_lo = -INFINITY, mx_hi = -INFINITY;
        const bool diag = (kt + 1) * BKV > pos0 + q0;  // only tiles that reach the panel can contain masked columns
#pragma unroll
        for (int j = 0; j < BKV / 8; j++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float v = sc[j][e];  // raw q.k; 1/sqrt(hd) is folded into the exponent below (the max is scale invariant)
                if (diag) {
                    const int tcol = kt * BKV + j * 8 + 2 * t4 + (e & 1);
                    if (tcol > ((e & 2) ? p_hi : p_lo)) v = -INFINITY;
                }
                sc[j][e] = v;
            }
            mx_lo = fmaxf(mx_lo, fmaxf(sc[j][0], sc[j][1]));
            mx_hi = fmaxf(mx_hi, fmaxf(sc[j][2], sc[j][3]));
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)), mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)), mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
        const float base_lo = mn_lo == -INFINITY ? 0.f : mn_lo, base_hi = mn_hi == -INFINITY ? 0.f : mn_hi;  // fully masked row so far
        const float c_lo = exp2f((m_lo - base_lo) * SC), c_hi = exp2f((m_hi - base_hi) * SC);
        m_lo = mn_lo, m_hi = mn_hi;
        float rs_lo = 0.f, rs_hi = 0.f;
        uint32_t pa[BKV / 16][4];
#pragma unroll
        for (int j = 0; j < BKV / 8; j++) {
            const float p0 = exp2f((sc[j][0] - base_lo) * SC), p1 = exp2f((sc[j][1] - base_lo) * SC);
            const float p2 = exp2f((sc[j][2] - base_hi) * SC), p3 = exp2f((sc[j][3] - base_hi) * SC);
            rs_lo += p0 + p1, rs_hi += p2 + p3;
            pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);  // A fragment of P for the 16 tokens of n-tiles (2i, 2i+1)
            pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        l_lo = l_lo * c_lo + rs_lo, l_hi = l_hi * c_hi + rs_hi;
#pragma unroll
        for (int j = 0; j < HD / 8; j++) o[j][0] *= c_lo, o[j][1] *= c_lo, o[j][2] *= c_hi, o[j][3] *= c_hi;
        // ---- O += P V ----
#pragma unroll
        for (int kc_ = 0; kc_ < BKV / 16; kc_++) {
#pragma unroll
            for (int jp = 0; jp < HD / 16; jp++) {
                uint32_t b[4];
                ldsm_x4_t(b, vb + sw(kc_ * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, jp * 2 + (lane >> 4)));
                mma_bf16(o[2 * jp], pa[kc_], b[0], b[1]);
                mma_bf16(o[2 * jp + 1], pa[kc_], b[2], b[3]);
            }
        }
        __syncthreads();  // the other buffer is refilled at the top of the next iteration
    }
    // ---- normalise and store ----
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1), l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1), l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
    for (int j = 0; j < HD / 8; j++) {
        const int d = j * 8 + 2 * t4;
        if (r_lo < M) *reinterpret_cast<uint32_t*>(out + (size_t)r_lo * q_dim + (size_t)h * HD + d) = pack_bf16x2(o[j][0] * i_lo, o[j][1] * i_lo);
        if (r_hi < M) *reinterpret_cast<uint32_t*>(out + (size_t)r_hi * q_dim + (size_t)h * HD + d) = pack_bf16x2(o[j][2] * i_hi, o[j][3] * i_hi);
    }
}

// ---- batched decode: the `group` query heads that share a kv head are processed TOGETHER on the tensor cores --------------------------
// With many sequences per step the per-head kernels above read every cached K / V row once per QUERY head (8x for Qwen3-32B) and
// need thousands of latency-bound CTAs.  Here a CTA owns one (sequence, kv head): the group's query rows (<= 16, zero padded) are the
// A operand of mma.sync m16n8k16, the cached rows stream once through a cp.async double buffer in 64-token tiles (each warp takes 16
// tokens of a tile), and softmax / P.V follow the flash recipe of the prefill kernel.  The warps' partial (max, sum, acc) are merged
// through shared memory; long contexts are split over CTAs and merged by kf_attn_combine_kernel.  q must already be normalised +
// rotated and the new K / V rows appended (kf_qknorm_rope_kvappend).
template <int HD>
__global__ void __launch_bounds__(128) kf_attn_gqa_kernel(uint16_t* __restrict__ out, float* __restrict__ ws, const uint16_t* __restrict__ q,
                                                          const uint16_t* __restrict__ kc, const uint16_t* __restrict__ vc,
                                                          const int32_t* __restrict__ pos_dev, int n_head, int n_kv, int nsplit, float sqrt_hd,
                                                          size_t seq_stride) {
    constexpr int CH = HD / 8;
    constexpr int TILE_BYTES = kPfKV * HD * 2;
    extern __shared__ __align__(16) uint8_t gq_smem[];
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(gq_smem);  // 16 x HD
    const uint32_t sK = sQ + 16 * HD * 2, sV = sK + 2 * TILE_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int kvh = blockIdx.x, m = blockIdx.y, split = blockIdx.z;
    const int group = n_head / n_kv, kv_dim = n_kv * HD;
    const int len = pos_dev[m] + 1;
    const int t0 = (int)(((long long)split * len) / nsplit), t1 = (int)(((long long)(split + 1) * len) / nsplit);
    const int ntiles = (t1 - t0 + kPfKV - 1) / kPfKV;
    auto sw = [](int row, int c) { return (uint32_t)((row * CH + (c ^ (row & 7))) * 16); };
    const uint16_t* kbase = kc + (size_t)m * seq_stride + (size_t)kvh * HD;
    const uint16_t* vbase = vc + (size_t)m * seq_stride + (size_t)kvh * HD;
    for (int i = tid; i < 16 * CH; i += 128) {
        const int r = i / CH, c = i % CH;
        const bool ok = r < group;
        cp16(sQ + sw(r, c), q + ((size_t)m * n_head + (size_t)kvh * group + (ok ? r : 0)) * HD + c * 8, ok);
    }
    auto load_kv = [&](int kt, int buf) {
        for (int i = tid; i < kPfKV * CH; i += 128) {
            const int r = i / CH, c = i % CH;
            const int t = min(t0 + kt * kPfKV + r, t1 - 1);  // the tail of the last tile is masked below
            cp16(sK + buf * TILE_BYTES + sw(r, c), kbase + (size_t)t * kv_dim + c * 8, true);
            cp16(sV + buf * TILE_BYTES + sw(r, c), vbase + (size_t)t * kv_dim + c * 8, true);
        }
    };
    if (ntiles > 0) load_kv(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t qa[HD / 16][4];
    float o[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    const float LOG2E = 1.4426950408889634f;
    for (int kt = 0; kt < ntiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) load_kv(kt + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int kc_ = 0; kc_ < HD / 16; kc_++) ldsm_x4(qa[kc_], sQ + sw(lane & 15, kc_ * 2 + (lane >> 4)));
        }
        const uint32_t kb = sK + buf * TILE_BYTES, vb = sV + buf * TILE_BYTES;
        const int tok0 = t0 + kt * kPfKV + warp * 16;  // this warp's 16 tokens of the tile
        if (tok0 < t1) {
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 2; j++) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
            for (int kc_ = 0; kc_ < HD / 16; kc_++) {
                uint32_t b[4];
                ldsm_x4(b, kb + sw(warp * 16 + (lane & 7) + (lane >> 4) * 8, kc_ * 2 + ((lane >> 3) & 1)));
                mma_bf16(sc[0], qa[kc_], b[0], b[1]);
                mma_bf16(sc[1], qa[kc_], b[2], b[3]);
            }
            float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2; j++) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int tcol = tok0 + j * 8 + 2 * t4 + (e & 1);
                    sc[j][e]       = tcol < t1 ? sc[j][e] / sqrt_hd : -INFINITY;  // the reference divides (operator.cuh:630)
                }
                mx_lo = fmaxf(mx_lo, fmaxf(sc[j][0], sc[j][1]));
                mx_hi = fmaxf(mx_hi, fmaxf(sc[j][2], sc[j][3]));
            }
            mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)), mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
            mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)), mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
            const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);  // finite: token tok0 is always valid
            const float c_lo = exp2f((m_lo - mn_lo) * LOG2E), c_hi = exp2f((m_hi - mn_hi) * LOG2E);
            m_lo = mn_lo, m_hi = mn_hi;
            float rs_lo = 0.f, rs_hi = 0.f;
            uint32_t pa[4];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float p0 = exp2f((sc[j][0] - mn_lo) * LOG2E), p1 = exp2f((sc[j][1] - mn_lo) * LOG2E);
                const float p2 = exp2f((sc[j][2] - mn_hi) * LOG2E), p3 = exp2f((sc[j][3] - mn_hi) * LOG2E);
                rs_lo += p0 + p1, rs_hi += p2 + p3;
                pa[j * 2 + 0] = pack_bf16x2(p0, p1), pa[j * 2 + 1] = pack_bf16x2(p2, p3);
            }
            l_lo = l_lo * c_lo + rs_lo, l_hi = l_hi * c_hi + rs_hi;
#pragma unroll
            for (int j = 0; j < HD / 8; j++) o[j][0] *= c_lo, o[j][1] *= c_lo, o[j][2] *= c_hi, o[j][3] *= c_hi;
#pragma unroll
            for (int jp = 0; jp < HD / 16; jp++) {
                uint32_t b[4];
                ldsm_x4_t(b, vb + sw(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, jp * 2 + (lane >> 4)));
                mma_bf16(o[2 * jp], pa, b[0], b[1]);
                mma_bf16(o[2 * jp + 1], pa, b[2], b[3]);
            }
        }
        __syncthreads();
    }
    // ---- merge the four warps (fixed order) through shared memory: the tile buffers are free now ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float* s_o = reinterpret_cast<float*>(gq_smem + 16 * HD * 2);  // [4][16][HD]
    float* s_m = s_o + 4 * 16 * HD;                                // [4][16]
    float* s_l = s_m + 64;
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1), l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1), l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
#pragma unroll
    for (int j = 0; j < HD / 8; j++) {
        float* r0 = s_o + ((size_t)warp * 16 + g) * HD + j * 8 + 2 * t4;
        float* r1 = r0 + 8 * HD;
        r0[0] = o[j][0], r0[1] = o[j][1], r1[0] = o[j][2], r1[1] = o[j][3];
    }
    if (t4 == 0) s_m[warp * 16 + g] = m_lo, s_m[warp * 16 + g + 8] = m_hi, s_l[warp * 16 + g] = l_lo, s_l[warp * 16 + g + 8] = l_hi;
    __syncthreads();
    for (int idx = tid; idx < group * HD; idx += 128) {
        const int r = idx / HD, c = idx % HD;
        float M_ = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; w++) M_ = fmaxf(M_, s_m[w * 16 + r]);
        float L_ = 0.f, acc = 0.f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const float sc_ = s_m[w * 16 + r] == -INFINITY ? 0.f : exp2f((s_m[w * 16 + r] - M_) * LOG2E);
            L_ += s_l[w * 16 + r] * sc_;
            acc = fmaf(s_o[((size_t)w * 16 + r) * HD + c], sc_, acc);
        }
        const int h = kvh * group + r;
        if (nsplit == 1) {
            out[((size_t)m * n_head + h) * HD + c] = f32_to_bf16_bits(acc / L_);
        } else {  // partial in the layout kf_attn_combine_kernel reads: [M][n_head][nsplit][HD + 2]
            float* wp = ws + (((size_t)m * n_head + h) * nsplit + split) * (HD + 2);
            wp[c] = acc;
            if (c == 0) wp[HD] = M_, wp[HD + 1] = L_;
        }
    }
}

// ---- QK-norm + RoPE + K/V append for big panels: one WARP per (token, head) (the block-per-head kernel of ops.cu needs 150 K tiny
// blocks for a 2048-token panel and is scheduling bound: 80 us against 15 us here).  Same arithmetic; the sum of squares is taken in
// warp-butterfly order as in the fused decode kernel.
template <int DPL>
__global__ void __launch_bounds__(256) kf_qknorm_rope_kv_warp_kernel(uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                                     const uint16_t* __restrict__ v, const uint16_t* __restrict__ qw,
                                                                     const uint16_t* __restrict__ kw, uint16_t* __restrict__ kcache,
                                                                     uint16_t* __restrict__ vcache, const float2* __restrict__ table,
                                                                     const int32_t* __restrict__ pos_dev, int M, int n_head, int n_kv, float eps,
                                                                     size_t seq_stride) {
    constexpr int HD = DPL * 32;
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int per_tok   = n_head + n_kv;
    if (wid >= (long long)M * per_tok) return;
    const int m = (int)(wid / per_tok), h = (int)(wid % per_tok);
    const int pos = pos_dev[m];
    const float2* cs_row = table + (size_t)pos * (HD / 2);
    const bool is_q = h < n_head;
    const int kvh   = h - n_head;
    const uint16_t* src = is_q ? q + ((size_t)m * n_head + h) * HD : k + ((size_t)m * n_kv + kvh) * HD;
    uint16_t* dst       = is_q ? q + ((size_t)m * n_head + h) * HD : kcache + (size_t)m * seq_stride + ((size_t)pos * n_kv + kvh) * HD;
    float o[DPL];
    norm_rope_row<DPL>(o, src, is_q ? qw : kw, cs_row, lane, eps);
    if constexpr (DPL == 4)
        *reinterpret_cast<uint2*>(dst + lane * 4) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    else
        *reinterpret_cast<uint32_t*>(dst + lane * 2) = pack_bf16x2(o[0], o[1]);
    if (!is_q) {  // the V row of this kv head
        const uint16_t* vs = v + ((size_t)m * n_kv + kvh) * HD + lane * DPL;
        uint16_t* vd       = vcache + (size_t)m * seq_stride + ((size_t)pos * n_kv + kvh) * HD + lane * DPL;
        if constexpr (DPL == 4)
            *reinterpret_cast<uint2*>(vd) = *reinterpret_cast<const uint2*>(vs);
        else
            *reinterpret_cast<uint32_t*>(vd) = *reinterpret_cast<const uint32_t*>(vs);
    }
}
}  // namespace

// ROPE::cuInfer (rope.cu:645-672) + attention_qk / softmax / attention_v (operator.cuh:573-668) of SelfAttention::cuInfer (QKV.cu:660-674)
// in ONE launch.  Precondition: the M tokens belong to M different sequences (seq_stride = per-sequence cache stride) or M == 1.
extern "C" int kf_qkv_attention(kf_ctx* ctx, void* out, const void* q, const void* k, const void* v, const void* qw, const void* kw, void* kc,
                                void* vc, const void* table, const int32_t* pos_dev, int M, int n_head, int n_kv, int hd, int max_seq,
                                float eps, size_t seq_stride, int max_pos_hint) {
    if (!ctx || !out || !q || !k || !v || !kc || !vc || !table || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (hd == 128 || hd == 64) && n_head % n_kv == 0 && M >= 1 && max_seq >= 1, "head_dim 64/128, GQA");
    KF_REQUIRE(ctx, M == 1 || seq_stride > 0, "the fused path needs one sequence per token (use kf_qknorm_rope_kvappend + kf_attn_decode for panels)");
#ifdef KF_DEBUG_KNOBS
    if (ctx->debug_skip & 1) return KF_OK;
#endif
    int nsplit = ctx->attn_split;
    const int len_hint = std::max(1, std::min(max_seq, max_pos_hint + 1));
    // contexts up to 2K tokens: every cached row of a warp's first pass is in flight before the dependency wait and the slices of a
    // head are merged inside a cluster.  Measured on B200 (32B decode, ctx 512; profiles/r01_attn_sweep.txt): what matters is that the
    // whole grid is resident in ONE wave (2 CTAs of 8 warps per SM) -- 4 slices x 8 warps x 8 rows beats 8 x 4 x 16 and the global-
    // workspace path; an 8-CTA cluster of 8 warps needs two waves and loses 20%.
    const int warps = ctx->attn_warps > 0 ? std::min(ctx->attn_warps, kClusterWarpsMax) : kClusterWarpsMax;
    if ((nsplit <= 0 && len_hint <= 4 * kMaxCluster * warps * kWarpTok) || (nsplit > 0 && nsplit <= kMaxCluster)) {
        if (nsplit <= 0) {
            const int one_wave = std::max(1, (2 * ctx->sm_count) / (n_head * M));
            nsplit = std::max(1, std::min(std::min(kMaxCluster, one_wave), (len_hint + 2 * warps * kWarpTok - 1) / (2 * warps * kWarpTok)));
        }
        dim3 grid(n_head, M, nsplit);
        const float sq = sqrtf((float)hd);
        if (hd == 128)
            KF_CUDA(ctx, launch_attn_cluster<4>(ctx, grid, nsplit, warps, (uint16_t*)out, (const uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v,
                                                (const uint16_t*)qw, (const uint16_t*)kw, (uint16_t*)kc, (uint16_t*)vc, (const float2*)table, pos_dev,
                                                n_head, n_kv, sq, eps, seq_stride));
        else
            KF_CUDA(ctx, launch_attn_cluster<2>(ctx, grid, nsplit, warps, (uint16_t*)out, (const uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v,
                                                (const uint16_t*)qw, (const uint16_t*)kw, (uint16_t*)kc, (uint16_t*)vc, (const float2*)table, pos_dev,
                                                n_head, n_kv, sq, eps, seq_stride));
        KF_LAUNCH_CHECK(ctx);
        return KF_OK;
    }
    if (nsplit <= 0) {
        const int len = len_hint;
        nsplit        = (4 * ctx->sm_count + n_head * M - 1) / (n_head * M);
        nsplit        = std::min(nsplit, std::max(1, len / (kAttnWarps * 2 * kTokBatch)));  // two batches per warp measured best
        nsplit        = std::max(1, std::min(nsplit, 64));
    }
    float* ws = nullptr;
    if (nsplit > 1) {
        int rc = kf_ensure_attn_ws(ctx, (size_t)M * n_head * nsplit * (hd + 2) * sizeof(float));
        if (!rc) rc = kf_ensure_attn_cnt(ctx, M * n_head);
        if (rc) return rc;
        ws = ctx->attn_ws;
    }
    dim3 grid(n_head, M, nsplit);
    const float sq = sqrtf((float)hd);
    if (hd == 128)
        KF_CUDA(ctx, kf_launch_pdl(ctx, kf_attn_fused_kernel<4>, grid, dim3(kAttnWarps * 32), 0, (uint16_t*)out, ws, ctx->attn_cnt, (const uint16_t*)q,
                                   (const uint16_t*)k, (const uint16_t*)v, (const uint16_t*)qw, (const uint16_t*)kw, (uint16_t*)kc, (uint16_t*)vc,
                                   (const float2*)table, pos_dev, n_head, n_kv, nsplit, sq, eps, seq_stride));
    else
        KF_CUDA(ctx, kf_launch_pdl(ctx, kf_attn_fused_kernel<2>, grid, dim3(kAttnWarps * 32), 0, (uint16_t*)out, ws, ctx->attn_cnt, (const uint16_t*)q,
                                   (const uint16_t*)k, (const uint16_t*)v, (const uint16_t*)qw, (const uint16_t*)kw, (uint16_t*)kc, (uint16_t*)vc,
                                   (const float2*)table, pos_dev, n_head, n_kv, nsplit, sq, eps, seq_stride));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

extern "C" int kf_attn_decode(kf_ctx* ctx, void* out, const void* q, const void* kc, const void* vc, const int32_t* pos_dev, int M, int n_head,
                              int n_kv, int hd, int max_seq, int max_pos_hint, size_t seq_stride) {
    if (!ctx || !out || !q || !kc || !vc || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (hd == 128 || hd == 64) && n_head % n_kv == 0 && M >= 1 && max_seq >= 1, "head_dim 64/128, GQA");
    // enough CTAs to cover the SMs, at least ~32 tokens per warp-slice
    int nsplit = ctx->attn_split;
    if (nsplit <= 0) {
        const int len = std::max(1, std::min(max_seq, max_pos_hint + 1));
        nsplit        = (4 * ctx->sm_count + n_head * M - 1) / (n_head * M);
        nsplit        = std::min(nsplit, std::max(1, len / (kAttnWarps * 2 * kTokBatch)));  // two batches per warp measured best
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

// Causal attention of a prefill panel: the M query rows q_dev[m] (already normalised + rotated, kf_qknorm_rope_kvappend) sit at the
// consecutive positions pos_dev[0] + m of ONE sequence whose K / V rows [0, pos_dev[0] + M) are in the cache layer.
extern "C" int kf_attn_prefill(kf_ctx* ctx, void* out, const void* q, const void* kc, const void* vc, const int32_t* pos_dev, int M, int n_head,
                               int n_kv, int hd, int max_seq) {
    if (!ctx || !out || !q || !kc || !vc || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (hd == 128 || hd == 64) && n_head % n_kv == 0 && M >= 1 && max_seq >= 1, "head_dim 64/128, GQA");
    const size_t smem = (size_t)(kPfQ + 4 * KF_PF_BKV) * hd * 2;
    dim3 grid((M + kPfQ - 1) / kPfQ, n_head);
    const float sq = sqrtf((float)hd);
    if (hd == 128) {
        static bool set[kf_ctx::kMaxDevices] = {};  // function attributes are per device
        if (!set[ctx->device]) {
            KF_CUDA(ctx, cudaFuncSetAttribute(kf_attn_prefill_kernel<128, KF_PF_BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set[ctx->device] = true;
        }
        kf_attn_prefill_kernel<128, KF_PF_BKV><<<grid, 128, smem, ctx->stream>>>((uint16_t*)out, (const uint16_t*)q, (const uint16_t*)kc, (const uint16_t*)vc,
                                                                      pos_dev, M, n_head, n_kv, max_seq, sq);
    } else {
        static bool set[kf_ctx::kMaxDevices] = {};  // function attributes are per device
        if (!set[ctx->device]) {
            KF_CUDA(ctx, cudaFuncSetAttribute(kf_attn_prefill_kernel<64, KF_PF_BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set[ctx->device] = true;
        }
        kf_attn_prefill_kernel<64, KF_PF_BKV><<<grid, 128, smem, ctx->stream>>>((uint16_t*)out, (const uint16_t*)q, (const uint16_t*)kc, (const uint16_t*)vc,
                                                                     pos_dev, M, n_head, n_kv, max_seq, sq);
    }
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

// Decode attention for MANY sequences per step (one token each): the query heads of a kv group share every cached row through the
// tensor cores (kf_attn_gqa_kernel).  Same contract as kf_attn_decode; q normalised + rotated, K / V of the current position appended.
extern "C" int kf_attn_decode_gqa(kf_ctx* ctx, void* out, const void* q, const void* kc, const void* vc, const int32_t* pos_dev, int M, int n_head,
                                  int n_kv, int hd, int max_seq, int max_pos_hint, size_t seq_stride) {
    if (!ctx || !out || !q || !kc || !vc || !pos_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (hd == 128 || hd == 64) && n_head % n_kv == 0 && n_head / n_kv <= 16 && M >= 1 && max_seq >= 1, "head_dim 64/128, group <= 16");
    KF_REQUIRE(ctx, M == 1 || seq_stride > 0, "one sequence per token");
    const int len = std::max(1, std::min(max_seq, max_pos_hint + 1));
    int nsplit    = ctx->attn_split;
    if (nsplit <= 0) {
        nsplit = (3 * ctx->sm_count + n_kv * M - 1) / (n_kv * M);
        nsplit = std::max(1, std::min(std::min(nsplit, 24), len / (2 * kPfKV)));  // measured at ctx 4096 / 8192: 12-24 slices equal, 32+ slower
    }
    float* ws = nullptr;
    if (nsplit > 1) {
        int rc = kf_ensure_attn_ws(ctx, (size_t)M * n_head * nsplit * (hd + 2) * sizeof(float));
        if (rc) return rc;
        ws = ctx->attn_ws;
    }
    const size_t smem = std::max((size_t)(16 + 4 * kPfKV) * hd * 2, (size_t)16 * hd * 2 + (size_t)4 * 16 * hd * 4 + 512);
    dim3 grid(n_kv, M, nsplit);
    const float sq = sqrtf((float)hd);
    if (hd == 128) {
        static bool set[kf_ctx::kMaxDevices] = {};  // function attributes are per device
        if (!set[ctx->device]) {
            KF_CUDA(ctx, cudaFuncSetAttribute(kf_attn_gqa_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set[ctx->device] = true;
        }
        kf_attn_gqa_kernel<128><<<grid, 128, smem, ctx->stream>>>((uint16_t*)out, ws, (const uint16_t*)q, (const uint16_t*)kc, (const uint16_t*)vc,
                                                                  pos_dev, n_head, n_kv, nsplit, sq, seq_stride);
    } else {
        static bool set[kf_ctx::kMaxDevices] = {};  // function attributes are per device
        if (!set[ctx->device]) {
            KF_CUDA(ctx, cudaFuncSetAttribute(kf_attn_gqa_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set[ctx->device] = true;
        }
        kf_attn_gqa_kernel<64><<<grid, 128, smem, ctx->stream>>>((uint16_t*)out, ws, (const uint16_t*)q, (const uint16_t*)kc, (const uint16_t*)vc,
                                                                 pos_dev, n_head, n_kv, nsplit, sq, seq_stride);
    }
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

// warp-per-head variant of kf_qknorm_rope_kvappend for big panels (called from ops.cu); hd 64 / 128 only
int kf_qknorm_rope_kv_warp(kf_ctx* ctx, void* q, const void* k, const void* v, const void* qw, const void* kw, void* kcache, void* vcache,
                           const void* table, const int32_t* pos_dev, int M, int n_head, int n_kv, int hd, float eps, size_t seq_stride) {
    const long long warps = (long long)M * (n_head + n_kv);
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    if (hd == 128)
        kf_qknorm_rope_kv_warp_kernel<4><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v, (const uint16_t*)qw,
                                                                          (const uint16_t*)kw, (uint16_t*)kcache, (uint16_t*)vcache, (const float2*)table,
                                                                          pos_dev, M, n_head, n_kv, eps, seq_stride);
    else
        kf_qknorm_rope_kv_warp_kernel<2><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v, (const uint16_t*)qw,
                                                                          (const uint16_t*)kw, (uint16_t*)kcache, (uint16_t*)vcache, (const float2*)table,
                                                                          pos_dev, M, n_head, n_kv, eps, seq_stride);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
