// sampler.cu -- temperature / top-k / top-p sampling ON THE DEVICE, one CTA per sequence: the step right after the logits
// (GeneratOnPrompt::Sample, reference src/Manifold/GoPT.cpp:614-630, with LogitsInfo::TopK :632-640, UpdateLogits :751-766, TopP :729-748,
// Qu_FlipCoin :768-786 and the xorshift64* generator :594-600).  The reference copies the [vocab] logits to the host every token and samples
// there; here only the 4-byte token leaves the GPU (or nothing at all inside the device-resident decode loop).
//
//   1. candidates  selection 0: the top_k largest logits, ties to the lower index (what TOPK_heap::Select is meant to do);
//                  selection 1: what TOPK_heap::Select (GoPT.cpp:667-700) actually keeps -- its std::priority_queue<int> orders INDICES, so
//                  the running "smallest kept" is always the newest index: indices 0 .. k-2 plus the first arg-max of the rest;
//                  both sorted by (logit descending, index ascending);
//   2. p_i = expf((l_i - l_max) / T), normalised by their sum in sorted order (sequential fp32, like the host loop);
//   3. top-p: keep the shortest prefix whose cumulative probability exceeds top_p;
//   4. coin = random_f32(state) * sum(kept p); first i with coin < cdf_i (else the last kept).
// Steps 2-4 run on one thread in the reference's order of operations, so that the CPU port (oracle kfo_sample) and the kernel agree
// except where expf's last ulp moves a cdf boundary across the coin.
#include <math.h>

#include "kf_common.cuh"

namespace {
constexpr int kSampThreads = 1024;
constexpr int kMaxTopK     = 1024;

__device__ __forceinline__ uint32_t order_key(uint16_t b) {  // bf16 bits -> 16-bit key, increasing with the value (-0 < +0 adjacent: equal as floats
    return (b & 0x8000u) ? (uint32_t)(uint16_t)~b : (uint32_t)(b | 0x8000u);  // is handled by comparing floats wherever ties matter)
}
__device__ __forceinline__ uint32_t xorshift_u32(unsigned long long* s) {  // GoPT.cpp:594-599
    *s ^= *s >> 12;
    *s ^= *s << 25;
    *s ^= *s >> 27;
    return (uint32_t)((*s * 0x2545F4914F6CDD1Dull) >> 32);
}

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = the sum
__device__ int block_exscan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(kSampThreads) kf_sample_kernel(int32_t* __restrict__ out, const uint16_t* __restrict__ logits, int vocab, float temperature,
                                                                 int top_k, float top_p, unsigned long long* __restrict__ rng, int selection) {
    __shared__ int s_hist[256];
    __shared__ int s_warp[33];
    __shared__ float s_val[kMaxTopK];
    __shared__ int s_idx[kMaxTopK];
    __shared__ int s_bin, s_above;
    const int tid       = threadIdx.x;
    const uint16_t* row = logits + (size_t)blockIdx.x * vocab;
    const int chunk     = (vocab + kSampThreads - 1) / kSampThreads;  // contiguous indices per thread: tie ranks follow the index order
    const int i0 = tid * chunk, i1 = min(vocab, i0 + chunk);
    const int k  = top_k;

    if (selection == 1) {
        // the reference's heap as it behaves: indices 0 .. k-2, plus the first maximum over [k-1, vocab)
        float best = -INFINITY;
        int bi     = 0x7fffffff;
        for (int i = max(i0, k - 1); i < i1; i++) {
            const float v = bf16_bits_to_f32(row[i]);
            if (v > best) best = v, bi = i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi   = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
        }
        __shared__ float s_bv[32];
        __shared__ int s_bi[32];
        if ((tid & 31) == 0) s_bv[tid >> 5] = best, s_bi[tid >> 5] = bi;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 32; w++)
                if (s_bv[w] > best || (s_bv[w] == best && s_bi[w] < bi)) best = s_bv[w], bi = s_bi[w];
            s_val[k - 1] = best, s_idx[k - 1] = bi;
        }
        if (tid < k - 1) s_val[tid] = bf16_bits_to_f32(row[tid]), s_idx[tid] = tid;
    } else {
        // ---- radix select of the k-th largest 16-bit key: high byte, then low byte inside that bin ---------------------------------
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int i = i0; i < i1; i++) atomicAdd(&s_hist[order_key(row[i]) >> 8], 1);
        __syncthreads();
        if (tid == 0) {
            int above = 0, b = 255;
            for (; b > 0 && above + s_hist[b] < k; b--) above += s_hist[b];
            s_bin = b, s_above = above;
        }
        __syncthreads();
        const int hb = s_bin;
        int above    = s_above;
        __syncthreads();
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int i = i0; i < i1; i++) {
            const uint32_t key = order_key(row[i]);
            if ((int)(key >> 8) == hb) atomicAdd(&s_hist[key & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int a = above, b = 255;
            for (; b > 0 && a + s_hist[b] < k; b--) a += s_hist[b];
            s_bin = b, s_above = a;
        }
        __syncthreads();
        const uint32_t thr = ((uint32_t)hb << 8) | (uint32_t)s_bin;  // the k-th largest key
        const int need     = k - s_above;                            // how many of the elements AT the threshold are kept: the lowest indices
        // ---- collect: everything above the threshold, and the first `need` at it (rank by index through a block scan) ---------------
        int n_above = 0, n_tie = 0;
        for (int i = i0; i < i1; i++) {
            const uint32_t key = order_key(row[i]);
            n_above += key > thr, n_tie += key == thr;
        }
        int tot_above, tot_tie;
        int off_above = block_exscan(n_above, s_warp, &tot_above);
        int off_tie   = block_exscan(n_tie, s_warp, &tot_tie);
        for (int i = i0; i < i1; i++) {
            const uint16_t b    = row[i];
            const uint32_t key = order_key(b);
            if (key > thr) {
                s_val[off_above] = bf16_bits_to_f32(b), s_idx[off_above] = i, off_above++;
            } else if (key == thr) {
                if (off_tie < need) s_val[tot_above + off_tie] = bf16_bits_to_f32(b), s_idx[tot_above + off_tie] = i;
                off_tie++;
            }
        }
    }
    for (int i = k + tid; i < kMaxTopK; i += kSampThreads) s_val[i] = -INFINITY, s_idx[i] = 0x7fffffff;
    __syncthreads();
    // ---- bitonic sort of the 1024 slots by (value descending, index ascending) --------------------------------------------------------
    for (int size = 2; size <= kMaxTopK; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int j = tid ^ stride;
            if (j > tid) {
                const float va = s_val[tid], vb = s_val[j];
                const int ia = s_idx[tid], ib = s_idx[j];
                const bool a_first = va > vb || (va == vb && ia < ib);  // a belongs before b in the final order
                const bool desc    = (tid & size) == 0;
                if (desc ? !a_first : a_first) s_val[tid] = vb, s_val[j] = va, s_idx[tid] = ib, s_idx[j] = ia;
            }
            __syncthreads();
        }
    // ---- softmax with temperature over the candidates, top-p, coin: one thread, the host loop's order of operations ---------------------
    if (tid == 0) {
        const float mx = s_val[0];
        float sum      = 0.f;
        for (int i = 0; i < k; i++) {
            const float p = expf((s_val[i] - mx) / temperature);
            s_val[i]      = p;
            sum += p;
        }
        for (int i = 0; i < k; i++) s_val[i] /= sum;
        int n_pick = k;
        if (top_p < 1.0f) {
            float cum = 0.f;
            int last  = k - 1;
            for (int i = 0; i < k; i++) {
                cum += s_val[i];
                if (cum > top_p) {
                    last = i;
                    break;
                }
            }
            n_pick = last + 1;
        }
        float psum = 0.f;
        for (int i = 0; i < n_pick; i++) psum += s_val[i];
        unsigned long long st = rng[blockIdx.x];
        const float coin      = (float)(xorshift_u32(&st) >> 8) / 16777216.0f * psum;
        rng[blockIdx.x]       = st;
        int q     = s_idx[n_pick - 1];
        float cdf = 0.f;
        for (int i = 0; i < n_pick; i++) {
            cdf += s_val[i];
            if (coin < cdf) {
                q = s_idx[i];
                break;
            }
        }
        out[blockIdx.x] = q;
    }
}
}  // namespace

// next[m] = a token drawn from row m of the logits; rng_state_dev: one 64-bit xorshift64* state per row (seeded by the caller, advanced here).
// temperature == 0 or top_k == 1: greedy (kf_argmax).  selection: 0 = true top-k, 1 = the reference's TOPK_heap::Select as it behaves.
extern "C" int kf_sample(kf_ctx* ctx, int32_t* next_dev, const void* logits_bf16_dev, int M, int vocab, float temperature, int top_k, float top_p,
                         uint64_t* rng_state_dev, int selection) {
    if (!ctx || !next_dev || !logits_bf16_dev) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1 && vocab >= 2, "shape");
    if (temperature == 0.0f || top_k == 1) return kf_argmax(ctx, next_dev, logits_bf16_dev, M, vocab);
    KF_REQUIRE(ctx, rng_state_dev, "sampling needs the generator state");
    KF_REQUIRE(ctx, temperature > 0.0f && top_p > 0.0f && (selection == 0 || selection == 1), "temperature > 0, top_p > 0");
    if (top_k <= 0 || top_k > vocab) top_k = vocab;  // CHAT_SAMPLER: nCanTopK = min(top_k, n_vocab), GoPT.cpp:389
    KF_REQUIRE(ctx, top_k <= kMaxTopK && top_k < vocab / 2, "top_k <= 1024 candidates (and < vocab / 2, TOPK_heap::Select's own assert)");
    kf_sample_kernel<<<M, kSampThreads, 0, ctx->stream>>>(next_dev, (const uint16_t*)logits_bf16_dev, vocab, temperature, top_k, top_p,
                                                          (unsigned long long*)rng_state_dev, selection);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
