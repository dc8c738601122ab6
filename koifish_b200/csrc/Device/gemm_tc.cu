// gemm_tc.cu -- dequant-fused tensor-core GEMM on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), persistent:
//   Y[M][N] = X[M][K] . deq(W[N][K])^T       tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), accumulators in TMEM
//
// Replaces the same reference pair as gemv.cu: GTensor::GetDataX (whole-matrix dequant to a bf16 scratch,
// src/Device/CUDA/kernel/quantizer.cu:249-392) + CU_mm_blasLt (src/Device/CUDA/kernel/gemm.cu:93-214) behind SLP::Forw
// (src/Device/CUDA/NeuronFuse.cu:305-381).  The packed weights cross HBM once per token panel and are expanded on chip.
//
// Work item = [128 weight rows] x [BN tokens] x [one K range (split-K)].  The weights are the A operand (UMMA_M = 128), the tokens
// the B operand (UMMA_N = BN), BK = 64.  One CTA per SM walks the items round-robin; its 15 warps are specialised:
//   * warp 9  (1 thread): TMA of the PACKED weight bytes, 64 B per row per stage, into a raw ring (SWIZZLE_64B, conflict-free reads);
//   * warp 10 (1 thread): TMA of the activation tile (already permuted, see below) into the B ring (SWIZZLE_128B, K-major);
//   * warps 0-7 (producers, thread = weight row x 32-weight slot): raw bytes -> the reference's bit-exact bf16 weights (the same
//     two-rounding arithmetic as gemv.cu) -> tcgen05.st into the A ring that lives in TENSOR MEMORY (lane = row, 2 weights / column);
//     the expanded weights never touch shared memory, which stays free for the B operand;
//   * warp 8  (1 thread): tcgen05.mma [D], [A in TMEM], B-descriptor; tcgen05.commit frees the stage / publishes the accumulator;
//   * warps 11-14 (epilogue): tcgen05.ld the accumulator (double-buffered, so the next item's main loop overlaps), then either the
//     final store (bf16 / +residual / fp32) or, with split-K, an fp32 partial + arrival counter; the last CTA to arrive reduces the
//     partials in fixed order (deterministic) and stores.
// bf16 weights need no expansion: warp 9 TMA-loads them straight into a swizzled A tile in shared memory (classic SS MMA).
// The k order inside every 32-wide slot is the extraction-friendly permutation of gemv.cu; kf_tc_prepare_x permutes the activations
// identically, so the contraction is unchanged.  Every barrier wait is bounded and traps instead of hanging the GPU.
#include <cuda.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <tuple>

#include "kf_common.cuh"

// tuning experiments (bit 0: no expansion, 1: no TMEM store, 2: no scale / zero fetch in the loop ...): debug builds only
// (KF_NVCC_DEFS="-DKF_DEBUG_KNOBS -DKF_TC_EXP=n"); a release build compiles them out
#if !defined(KF_DEBUG_KNOBS) || !defined(KF_TC_EXP)
#undef KF_TC_EXP
#define KF_TC_EXP 0
#endif

namespace {

enum { TF_BF16 = 0, TF_F8 = 1, TF_Q4 = 2, TF_Q2 = 3, TF_Q1 = 4 };
enum { TM_PLAIN = 0, TM_AFFINE = 1, TM_AFFINE_SYM = 2, TM_SCALE = 3, TM_AFFINE_FMA = 4 };  // _FMA: one bf16 rounding (ctx deq_fma, default), see kf_common.cuh

constexpr int BM = 128;    // weight rows per item (UMMA M)
constexpr int BK = 64;     // k per stage: 64 bf16 = one 128-byte swizzle row of the B tile
#ifndef KF_TC_RAWB
#define KF_TC_RAWB 128
#endif
#ifndef KF_TC_RAWB_LOWBIT
#define KF_TC_RAWB_LOWBIT 64
#endif
// packed bytes per weight row per raw stage (64: SWIZZLE_64B, 128: SWIZZLE_128B).  A K split ends on a raw-stage boundary, and 128 bytes
// are 512 / 1024 weights of the 2- / 1-bit formats: too coarse to balance 148 SMs (1-bit 25600x5120 at 32 tokens: 71 us, 38 us with 64)
template <int FMT>
__host__ __device__ constexpr int raw_bytes() { return (FMT == 3 || FMT == 4) ? KF_TC_RAWB_LOWBIT : KF_TC_RAWB; }  // TF_Q2, TF_Q1
constexpr int kProducerWarps = 8, kMmaWarp = 8, kRawWarp = 9, kXWarp = 10, kEpiWarp0 = 11;
constexpr int kThreadsTC = 15 * 32;

constexpr int kMaxW = 3;  // weights of one launch (Q/K/V, gate/up): same type, same K, their row tiles share one item space
struct GemmParams {
    const uint16_t* zero[kMaxW];
    const uint16_t* step[kMaxW];
    void* y[kMaxW];            // bf16 [M][N_w] (or float when epilogue == 4)
    int Nw[kMaxW];             // rows of each weight
    int tile0[kMaxW + 1];      // first row tile of each weight in the launch's tile space (unused entries: INT_MAX)
    const uint16_t* residual;  // bf16 [M][N] or nullptr (single weight)
    float* ws;                 // split-K partials
    unsigned* cnt;             // split-K arrival counters (self-resetting)
    int M, N, K;
    int qbias, gshift, epilogue;
    int n_tiles, m_tiles, splits, n_items;
    uint32_t lop_mask, lop_magic;
};

template <int FMT>
struct Fmt {
    static constexpr int BITS  = FMT == TF_BF16 ? 16 : FMT == TF_F8 ? 8 : FMT == TF_Q4 ? 4 : FMT == TF_Q2 ? 2 : 1;
    static constexpr int SLOTB = 4 * BITS;                             // packed bytes of one 32-weight slot
    static constexpr int RAWB  = raw_bytes<FMT>();
    static constexpr int KBR   = FMT == TF_BF16 ? 2 : RAWB / (2 * SLOTB);  // k-blocks per raw stage (bf16: granularity of a K split)
};

// ---- k permutation inside a 32-wide slot (identical to gemv.cu's xperm) ------------------------------------------------------------
template <int FMT>
__host__ __device__ constexpr int tperm(int o) {
    const int u = o >> 3, e = o & 7;
    if (FMT == TF_Q4) return 8 * u + ((e & 1) ? 3 : 7) - (e >> 1);
    if (FMT == TF_Q2) return 16 * (u >> 1) + ((e & 1) ? 7 : 15) - (e >> 1) - 4 * (u & 1);
    if (FMT == TF_Q1) return ((e & 1) ? 15 : 31) - (e >> 1) - 4 * u;
    return o;
}
template <int FMT>
__global__ void __launch_bounds__(256) kf_permute_x_kernel(uint16_t* __restrict__ xp, const uint16_t* __restrict__ x, size_t n_slots) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 32-element slot per thread
    if (i >= n_slots) return;
    const uint4* src = reinterpret_cast<const uint4*>(x + i * 32);
    uint32_t s[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 v = __ldg(src + j);
        s[4 * j] = v.x, s[4 * j + 1] = v.y, s[4 * j + 2] = v.z, s[4 * j + 3] = v.w;
    }
    uint32_t o[16];
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int e0 = tperm<FMT>(2 * p), e1 = tperm<FMT>(2 * p + 1);
        o[p] = ((s[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu) | (((s[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu) << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(xp + i * 32);
#pragma unroll
    for (int j = 0; j < 4; j++) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}

// RMSNorm (the arithmetic and summation order of kf_rmsnorm_kernel, ops.cu: bit-identical output) written directly in the permuted
// k order: one launch instead of norm + permute in front of the Q/K/V and gate/up launches
template <int FMT>
__global__ void __launch_bounds__(256) kf_rmsnorm_permute_kernel(uint16_t* __restrict__ xp, const uint16_t* __restrict__ x, const uint16_t* __restrict__ w,
                                                                  int dim, float eps) {
    __shared__ float red[32];
    const uint16_t* xr = x + (size_t)blockIdx.x * dim;
    uint16_t* orow     = xp + (size_t)blockIdx.x * dim;
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
    for (int slot = threadIdx.x; slot * 32 < dim; slot += blockDim.x) {
        uint32_t n[16];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 v = *reinterpret_cast<const uint4*>(xr + slot * 32 + c * 8);
            const uint4 g = *reinterpret_cast<const uint4*>(w + slot * 32 + c * 8);
            const uint32_t q[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int j = 0; j < 4; j++) n[4 * c + j] = pack_bf16x2((bf16lo(q[j]) * s) * bf16lo(gw[j]), (bf16hi(q[j]) * s) * bf16hi(gw[j]));
        }
        uint32_t o[16];
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int e0 = tperm<FMT>(2 * p), e1 = tperm<FMT>(2 * p + 1);
            o[p] = ((n[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu) | (((n[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu) << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(orow + slot * 32);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
}

// ---- PTX helpers ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    return ok;
}
// bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __noinline__ void mbar_wait_slow(uint32_t a, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 24); it++)
        if (mbar_try(a, parity)) return;
    __trap();
}
__device__ __forceinline__ void mbar_wait_a(uint32_t a, uint32_t parity) {
    if (!mbar_try(a, parity)) mbar_wait_slow(a, parity);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_a(smem_u32(bar), parity); }
// for the warps that wait a long time (epilogue): back off between polls so the spinning does not take issue slots from the producers
__device__ __noinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); it++) {
        if (mbar_try(a, parity)) return;
        __nanosleep(200);
    }
    __trap();
}
__device__ __forceinline__ void lds128(uint32_t a, uint32_t* r) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void lds64(uint32_t a, uint32_t* r) {
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units          bits [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)   [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)           [46,48)
    d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B                 [61,64)
    return d;
}
// instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, M = 128, N = n (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// The MMA / commit helpers are executed by the whole (converged) warp; elect.sync picks one lane -- the same one every time, so all
// MMAs and the commits that track them come from a single thread -- and only that lane issues the instruction.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_addr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
                 "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ uint32_t and_or3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// one pair of codes -> bf16x2 weights, bit-exact w.r.t. the reference (see gemv.cu deq_pair)
template <int MODE>
__device__ __forceinline__ uint32_t tdeq(uint32_t reg, int shift, uint32_t step2, uint32_t zero2, uint32_t nb2, uint32_t bias2, uint32_t mask,
                                         uint32_t magic) {
    const uint32_t v = and_or3(reg >> shift, mask, magic);
    if (MODE == TM_AFFINE_FMA) {  // zero2 holds -zero: RN(step * k - zero), ONE rounding (fma.rn.bf16, as the reference built for sm_90+)
        __nv_bfloat162 k = __hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2));
        return bf162_as_u32(__hfma2(k, u32_as_bf162(step2), u32_as_bf162(zero2)));
    }
    if (MODE == TM_AFFINE) {
        __nv_bfloat162 p = __hfma2(u32_as_bf162(v), u32_as_bf162(step2), u32_as_bf162(nb2));
        return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
    }
    __nv_bfloat162 k = __hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2));
    __nv_bfloat162 p = __hmul2_rn(u32_as_bf162(step2), k);  // ternary / binary: exactly {-step, 0, step} / {0, step}
    if (MODE == TM_SCALE) return bf162_as_u32(p);
    return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
}
__device__ __forceinline__ uint32_t tf8_pair(uint32_t reg, uint32_t sel) {
    uint32_t h2 = __byte_perm(reg, 0u, sel);
    uint32_t t  = ((h2 >> 3) & 0x0FE00FE0u) | (h2 & 0x80008000u);
    return bf162_as_u32(__hmul2_rn(u32_as_bf162(t), u32_as_bf162(0x77807780u)));
}

// Expand the 32 weights of one slot (codes in `w`, memory order) into 16 bf16x2 registers in the permuted k order.
template <int FMT, int MODE>
__device__ __forceinline__ void expand_slot(uint32_t (&o)[16], const uint32_t* w, uint32_t step2, uint32_t zero2, uint32_t nb2, uint32_t bias2,
                                            uint32_t mask, uint32_t magic) {
    if constexpr (FMT == TF_Q4) {  // 4 registers, first codes in the LAST register, nibble c_j at bits [31-4j : 28-4j]
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = w[3 - u];
            o[4 * u + 0] = tdeq<MODE>(r, 0, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 1] = tdeq<MODE>(r, 4, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 2] = tdeq<MODE>(r, 8, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 3] = tdeq<MODE>(r, 12, step2, zero2, nb2, bias2, mask, magic);
        }
    } else if constexpr (FMT == TF_Q2) {  // 2 registers
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = w[1 - (u >> 1)];
            const int b      = 8 * (u & 1);
            o[4 * u + 0] = tdeq<MODE>(r, b + 0, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 1] = tdeq<MODE>(r, b + 2, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 2] = tdeq<MODE>(r, b + 4, step2, zero2, nb2, bias2, mask, magic);
            o[4 * u + 3] = tdeq<MODE>(r, b + 6, step2, zero2, nb2, bias2, mask, magic);
        }
    } else if constexpr (FMT == TF_Q1) {  // 1 register
#pragma unroll
        for (int p = 0; p < 16; p++) o[p] = tdeq<MODE>(w[0], p, step2, zero2, nb2, bias2, mask, magic);
    } else {  // F8: 8 registers, natural order
#pragma unroll
        for (int r = 0; r < 8; r++) o[2 * r] = tf8_pair(w[r], 0x1404u), o[2 * r + 1] = tf8_pair(w[r], 0x3424u);
    }
}

// byte offset of 32-weight slot q inside the 64-byte raw row of a stage (PackedQ.hpp:160-211: the first codes sit in word.high)
template <int FMT>
__device__ __forceinline__ int slot_offset(int q) {
    if (FMT == TF_Q4) return 16 * q;
    if (FMT == TF_Q2) return 16 * (q >> 1) + 8 * (1 - (q & 1));
    if (FMT == TF_Q1) return 16 * (q >> 2) + 12 - 4 * (q & 3);
    return 32 * q;  // F8
}

// Ring depths.  The three rings are decoupled because their latencies differ: the raw ring hides DRAM latency (bytes in flight per SM
// = RS x 8 KB must cover ~2 us x 44 GB/s), the B ring hides the L2 round trip of the (small, L2-resident) activation tiles, the A ring
// only the on-chip producer -> MMA hand-off.  Sized so that shared memory stays below ~210 KB and tensor memory within 512 columns.
template <int FMT, int BN>
struct Rings {
    static constexpr bool A_TMEM = FMT != TF_BF16;
    static constexpr int U       = BN <= 128 ? 2 : 1;  // k-blocks per ring stage: below 256 tokens the barrier traffic, not the MMA, paces a k-block
    static constexpr int NACC    = (A_TMEM && BN == 256) ? 1 : 2;
    static constexpr int A_BYTES = A_TMEM ? 0 : BM * 128;
    static constexpr int SUB     = A_BYTES + BN * 128;  // one k-block of a stage: [A tile (bf16 weights only)][B tile]
    static constexpr int STAGE   = U * SUB;
    static constexpr int SA      = !A_TMEM ? 0 : (BN <= 64 ? 12 : 8) / U;
    static constexpr int SB      = (A_TMEM ? (BN == 16 ? 32 : BN == 32 ? 16 : BN == 64 ? 10 : BN == 128 ? 6 : 4)
                                           : (BN == 16 ? 10 : BN == 32 ? 10 : BN == 64 ? 8 : BN == 128 ? 6 : 4)) / U;
    static constexpr int RAWB    = raw_bytes<FMT>();
    static constexpr int RS      = !A_TMEM ? 0 : (BN <= 64 ? 128 : BN == 128 ? 112 : 80) * 1024 / (BM * RAWB);
    static constexpr size_t SMEM = (size_t)SB * STAGE + (size_t)RS * BM * RAWB + 1024;
    static_assert(SMEM <= 216 * 1024, "shared memory budget");
};

struct Item {
    int n0, m0, z, tile;  // tile = tn * m_tiles + mt
    int wi;               // weight of the launch
    int kb0, kb1;         // k-block range
};
template <int KBR>
__device__ __forceinline__ Item decode_item(const GemmParams& p, int item) {
    Item it;
    const int mt = item % p.m_tiles;
    const int t  = item / p.m_tiles;
    it.z         = t % p.splits;
    const int tn = t / p.splits;
    it.wi = (tn >= p.tile0[1] ? 1 : 0) + (tn >= p.tile0[2] ? 1 : 0);
    it.n0 = (tn - p.tile0[it.wi]) * BM, it.m0 = mt, it.tile = tn * p.m_tiles + mt;  // m0 = token-tile index (the caller scales it by BN)
    const int nkb  = p.K / BK;
    const int nraw = (nkb + KBR - 1) / KBR;
    const int base = nraw / p.splits, rem = nraw % p.splits;
    const int r0 = it.z * base + min(it.z, rem), r1 = r0 + base + (it.z < rem ? 1 : 0);
    it.kb0 = r0 * KBR, it.kb1 = min(r1 * KBR, nkb);
    return it;
}

// Store NC consecutive tokens (the first `valid` of them exist) of weight row `grow`: bf16, bf16 + residual, or fp32.  The residual
// values are all fetched before the first store (y may alias the residual, element for element).
template <int NC>
__device__ __forceinline__ void store_cols(const GemmParams& p, void* y, int N, const float (&acc)[NC], int m_first, int valid, int grow) {
    const size_t idx0 = (size_t)m_first * N + grow;
    if (p.epilogue == 4) {
#pragma unroll
        for (int j = 0; j < NC; j++)
            if (j < valid) reinterpret_cast<float*>(y)[idx0 + (size_t)j * N] = acc[j];
        return;
    }
    uint16_t res[NC];
    if (p.epilogue == 1) {
#pragma unroll
        for (int j = 0; j < NC; j++) res[j] = j < valid ? p.residual[idx0 + (size_t)j * N] : (uint16_t)0;
    }
#pragma unroll
    for (int j = 0; j < NC; j++) {
        if (j < valid) {
            uint16_t b = f32_to_bf16_bits(acc[j]);  // the reference's GEMM output is bf16 (gemm.cu:124-126)
            if (p.epilogue == 1) b = f32_to_bf16_bits(bf16_bits_to_f32(res[j]) + bf16_bits_to_f32(b));
            reinterpret_cast<uint16_t*>(y)[idx0 + (size_t)j * N] = b;
        }
    }
}

template <int FMT, int MODE, int BN>
__global__ void __launch_bounds__(kThreadsTC, 1)
    kf_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                      const __grid_constant__ CUtensorMap tm_x, const GemmParams p) {
    using F                  = Fmt<FMT>;
    using C                  = Rings<FMT, BN>;
    constexpr bool A_TMEM    = C::A_TMEM;
    constexpr int KBR        = F::KBR;
    constexpr int SA         = C::SA;    // A ring (expanded weights, tensor memory, 32 columns per stage)
    constexpr int SB         = C::SB;    // B ring (activation tiles, shared memory; for bf16 weights the stage also holds the A tile)
    constexpr int RS         = C::RS;    // raw ring (packed weight bytes, shared memory)
    constexpr int NACC       = C::NACC;  // accumulator buffers in tensor memory
    constexpr int A_BYTES    = C::A_BYTES;
    constexpr int B_BYTES    = BN * 128;
    constexpr int STAGE      = C::STAGE;
    constexpr int SUB        = C::SUB;
    constexpr int U          = C::U;     // k-blocks per stage of the A and B rings
    constexpr int RAWB       = F::RAWB;
    constexpr int RAW_BYTES  = BM * RAWB;
    constexpr int A_COL0     = NACC * BN;  // TMEM: accumulators first, then the A ring
    constexpr int TMEM_NEED  = A_COL0 + SA * U * 32;
    constexpr int TMEM_COLS  = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
    static_assert(TMEM_NEED <= 512, "tensor memory has 512 columns");
    constexpr int SA1 = SA > 0 ? SA : 1, RS1 = RS > 0 ? RS : 1;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* raws  = tiles + (size_t)SB * STAGE;
    __shared__ uint64_t a_full[SA1], a_empty[SA1], b_full[SB], b_empty[SB], raw_full[RS1], raw_empty[RS1], tmem_full[NACC], tmem_empty[NACC];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int last_flag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < SA; s++) mbar_init(&a_full[s], kProducerWarps), mbar_init(&a_empty[s], 1);  // one arrival per producer WARP
        for (int s = 0; s < SB; s++) mbar_init(&b_full[s], A_TMEM ? 1 : 2), mbar_init(&b_empty[s], 1);
        for (int s = 0; s < RS; s++) mbar_init(&raw_full[s], 1), mbar_init(&raw_empty[s], kProducerWarps);
        for (int s = 0; s < NACC; s++) mbar_init(&tmem_full[s], 1), mbar_init(&tmem_empty[s], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    kf_grid_launch_dependents();

    if (warp < kProducerWarps) {
        if constexpr (A_TMEM) {
            // ============ producers: thread = (weight row, 32-weight slot of the k-block): raw bytes -> bf16 -> TMEM ============
            const int quad = warp & 3, slot = warp >> 2;
            const int row  = quad * 32 + lane;
            const int gpr  = (p.K >> 7) >> p.gshift;
            const uint32_t bias2 = pack_bf16x2((float)(128 + p.qbias), (float)(128 + p.qbias));
            // TMA swizzle of the raw stage: the 16-byte chunk index is XORed with address bits [7,9) (64-byte rows) / [7,10) (128-byte rows)
            const uint32_t rsw   = RAWB == 64 ? (uint32_t)((row >> 1) & 3) : (uint32_t)(row & 7);
            constexpr int UNIT = KBR >= 2 ? 2 : 1;  // k-blocks expanded together (independent work for the scheduler)
            constexpr int NW   = F::SLOTB / 4;      // registers of one packed slot
            const uint32_t raw_thread = smem_u32(raws) + (uint32_t)(row * RAWB);
            const uint32_t a_thread   = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(A_COL0 + slot * 16);
            const uint32_t full0 = smem_u32(&a_full[0]), empty0 = smem_u32(&a_empty[0]);
            const uint32_t rfull0 = smem_u32(&raw_full[0]), rempty0 = smem_u32(&raw_empty[0]);
            uint32_t s = 0, eph = 1, rs = 0, rph = 0;  // ring positions and the parities to wait for
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item w = decode_item<KBR>(p, item);
                const int grow       = min(w.n0 + row, p.Nw[w.wi] - 1);
                const uint16_t* zrow = MODE == TM_PLAIN ? nullptr : p.zero[w.wi] + (size_t)grow * gpr;
                const uint16_t* srow = MODE == TM_PLAIN ? nullptr : p.step[w.wi] + (size_t)grow * gpr;
                // scale / zero of the current group and of the next two (register queue; the loads run 1-2 groups ahead of their use)
                int gcur = 0;
                uint32_t zq0 = 0, sq0 = 0, zq1 = 0, sq1 = 0, zq2 = 0, sq2 = 0;
                uint32_t step2 = 0, zero2 = 0, nb2 = 0;
                if (MODE != TM_PLAIN) {
                    gcur = (w.kb0 >> 1) >> p.gshift;
                    zq0 = __ldg(zrow + gcur), sq0 = __ldg(srow + gcur);
                    const int g1 = min(gcur + 1, gpr - 1), g2 = min(gcur + 2, gpr - 1);
                    zq1 = __ldg(zrow + g1), sq1 = __ldg(srow + g1);
                    zq2 = __ldg(zrow + g2), sq2 = __ldg(srow + g2);
                    step2 = __byte_perm(sq0, 0u, 0x1010), zero2 = __byte_perm(zq0, 0u, 0x1010) ^ (MODE == TM_AFFINE_FMA ? 0x80008000u : 0u);
                    if (MODE == TM_AFFINE) nb2 = bf162_as_u32(__hmul2_rn(u32_as_bf162(step2), u32_as_bf162(0xC300C300u)));
                }
                const int r1 = (w.kb1 + KBR - 1) / KBR;
                for (int r = w.kb0 / KBR; r < r1; r++) {
                    // ---- this thread's slots of the raw stage: one per k-block ----
                    mbar_wait_a(rfull0 + rs * 8, rph);
                    const uint32_t src = raw_thread + rs * RAW_BYTES;
                    uint32_t wreg[KBR][NW];
#pragma unroll
                    for (int kin = 0; kin < KBR; kin++) {
                        const int boff = slot_offset<FMT>(kin * 2 + slot);
                        if constexpr (F::SLOTB == 32) {
                            lds128(src + ((((uint32_t)(boff >> 4) + 0) ^ rsw) << 4), &wreg[kin][0]);
                            lds128(src + ((((uint32_t)(boff >> 4) + 1) ^ rsw) << 4), &wreg[kin][4]);
                        } else if constexpr (F::SLOTB == 16) {
                            lds128(src + (((uint32_t)(boff >> 4) ^ rsw) << 4), &wreg[kin][0]);
                        } else if constexpr (F::SLOTB == 8) {
                            lds64(src + (((uint32_t)(boff >> 4) ^ rsw) << 4) + (boff & 15), &wreg[kin][0]);
                        } else {
                            wreg[kin][0] = lds32(src + (((uint32_t)(boff >> 4) ^ rsw) << 4) + (boff & 15));
                        }
                    }
                    // the bytes are in registers: hand the stage back to the loader.  One arrival per warp: 256 per-thread arrivals on
                    // one barrier word serialise in the shared-memory atomic unit and were the bottleneck of the whole kernel
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(rempty0 + rs * 8);
                    if (++rs == RS) rs = 0, rph ^= 1;
#pragma unroll
                    for (int kin = 0; kin < KBR; kin += UNIT) {
                        const int kb = r * KBR + kin;
                        if (KBR > 2 && kb >= w.kb1) break;  // ragged last raw stage (K is a multiple of 128: never splits a unit)
                        if (MODE != TM_PLAIN) {
                            const int gi = (kb >> 1) >> p.gshift;
                            if (gi != gcur) {  // rotate the queue, fetch two groups ahead
                                gcur = gi;
                                zq0 = zq1, sq0 = sq1, zq1 = zq2, sq1 = sq2;
                                const int g2 = min(gi + 2, gpr - 1);
#if !(KF_TC_EXP & 4)
                                zq2 = __ldg(zrow + g2), sq2 = __ldg(srow + g2);
#endif
                                step2 = __byte_perm(sq0, 0u, 0x1010), zero2 = __byte_perm(zq0, 0u, 0x1010) ^ (MODE == TM_AFFINE_FMA ? 0x80008000u : 0u);
                                if (MODE == TM_AFFINE) nb2 = bf162_as_u32(__hmul2_rn(u32_as_bf162(step2), u32_as_bf162(0xC300C300u)));
                            }
                        }
                        uint32_t o[UNIT][16];
#pragma unroll
                        for (int u = 0; u < UNIT; u++) {
#if KF_TC_EXP & 1
                            for (int i = 0; i < 16; i++) o[u][i] = wreg[kin + u][i % NW] ^ step2;
#else
                            expand_slot<FMT, MODE>(o[u], wreg[kin + u], step2, zero2, nb2, bias2, p.lop_mask, p.lop_magic);
#endif
                        }
                        uint32_t sfull[UNIT];
#pragma unroll
                        for (int u = 0; u < UNIT; u++) {
                            if (u % U == 0) {
                                mbar_wait_a(empty0 + s * 8, eph);
                                tc_fence_after();
                            }
#if KF_TC_EXP & 2
                            asm volatile("" ::"r"(o[u][0] ^ o[u][5] ^ o[u][10] ^ o[u][15]));
#else
                            tmem_st16(a_thread + s * (32 * U) + (u % U) * 32, o[u]);
#endif
                            sfull[u] = full0 + s * 8;
                            if (u % U == U - 1)
                                if (++s == SA) s = 0, eph ^= 1;
                        }
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int u = U - 1; u < UNIT; u += U) mbar_arrive_a(sfull[u]);
                        }
                    }
                }
            }
        }
    } else if (warp == kRawWarp) {
        if (lane == 0) {
            // ============ weight loader: packed bytes (or, for bf16, the A tile itself) by TMA ============
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item w = decode_item<KBR>(p, item);
                const CUtensorMap* tm_w = w.wi == 0 ? &tm_w0 : w.wi == 1 ? &tm_w1 : &tm_w2;
                if constexpr (A_TMEM) {
                    for (int r = w.kb0 / KBR; r * KBR < w.kb1; r++) {
                        const int rs = it % RS;
                        mbar_wait(&raw_empty[rs], ((it / RS) & 1) ^ 1);
#if KF_TC_EXP & 8
                        mbar_arrive(&raw_full[rs]);
#else
                        mbar_arrive_expect_tx(&raw_full[rs], RAW_BYTES);
                        tma_load_2d(raws + (size_t)rs * RAW_BYTES, tm_w, r * RAWB, w.n0, &raw_full[rs]);
#endif
                        it++;
                    }
                } else {
                    for (int kb = w.kb0; kb < w.kb1; kb += U) {
                        const int s = it % SB;
                        mbar_wait(&b_empty[s], ((it / SB) & 1) ^ 1);
                        mbar_arrive_expect_tx(&b_full[s], U * A_BYTES);
#pragma unroll
                        for (int u = 0; u < U; u++) tma_load_2d(tiles + (size_t)s * STAGE + u * SUB, tm_w, (kb + u) * BK, w.n0, &b_full[s]);
                        it++;
                    }
                }
            }
        }
    } else if (warp == kXWarp) {
        if (lane == 0) {
            // ============ activation loader: B tile by TMA (reads what the previous kernel wrote -> dependency wait first) ============
            kf_grid_dependency_wait();
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item w = decode_item<KBR>(p, item);
                for (int kb = w.kb0; kb < w.kb1; kb += U) {
                    const int s = it % SB;
                    mbar_wait(&b_empty[s], ((it / SB) & 1) ^ 1);
#if KF_TC_EXP & 16
                    mbar_arrive(&b_full[s]);
#else
                    mbar_arrive_expect_tx(&b_full[s], U * B_BYTES);
#pragma unroll
                    for (int u = 0; u < U; u++)
                        tma_load_2d(tiles + (size_t)s * STAGE + u * SUB + A_BYTES, &tm_x, (kb + u) * BK, w.m0 * BN, &b_full[s]);
#endif
                    it++;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ============ MMA issuer.  The whole warp walks the loop converged (ring state and descriptors stay in uniform registers);
        // one elected lane -- always the same one -- issues the MMAs and the commits (elect.sync inside the asm) ============
        constexpr uint32_t idesc = make_idesc_bf16(BN);
        uint32_t ait = 0;
        uint32_t sa = 0, aph = 0, sb = 0, bph = 0;  // ring positions and the parities to wait for
        const uint32_t bfull0 = smem_u32(&b_full[0]), bempty0 = smem_u32(&b_empty[0]);
        const uint32_t afull0 = smem_u32(&a_full[0]), aempty0 = smem_u32(&a_empty[0]);
        const uint32_t tiles0 = smem_u32(tiles);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const Item w  = decode_item<KBR>(p, item);
            const int buf = ait % NACC;
            mbar_wait(&tmem_empty[buf], ((ait / NACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
            uint32_t acc = 0;
            for (int kb = w.kb0; kb < w.kb1; kb += U) {
                {  // both polls are issued before either result is consumed (a try_wait takes ~90 cycles even when complete)
                    const uint32_t ab = bfull0 + sb * 8, aa = afull0 + (A_TMEM ? sa : 0) * 8;
                    const uint32_t okb = mbar_try(ab, bph), oka = A_TMEM ? mbar_try(aa, aph) : 1u;
                    if (!okb) mbar_wait_slow(ab, bph);
                    if (!oka) mbar_wait_slow(aa, aph);
                }
                tc_fence_after();
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t st_addr = tiles0 + sb * STAGE + u * SUB;
                    const uint64_t bdesc   = make_desc_sw128(st_addr + A_BYTES);
                    if constexpr (A_TMEM) {
                        const uint32_t ta = tmem_base + (uint32_t)(A_COL0 + sa * (32 * U) + u * 32);
#pragma unroll
                        for (int k = 0; k < ((KF_TC_EXP & 32) ? 0 : BK / 16); k++)  // 16 bf16 of K per MMA: +32 bytes in the swizzle row / +8 TMEM columns
                            umma_ts(tmem_d, ta + (uint32_t)(k * 8), bdesc + (uint64_t)(2 * k), idesc, (k > 0 || u > 0) ? 1u : acc);
                    } else {
                        const uint64_t adesc = make_desc_sw128(st_addr);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++)
                            umma_ss(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (k > 0 || u > 0) ? 1u : acc);
                    }
                }
                acc = 1;
                // the commits arrive when the MMAs above have finished reading the stages
                umma_commit_a(bempty0 + sb * 8);
                if (++sb == SB) sb = 0, bph ^= 1;
                if constexpr (A_TMEM) {
                    umma_commit_a(aempty0 + sa * 8);
                    if (++sa == SA) sa = 0, aph ^= 1;
                }
            }
            umma_commit_a(smem_u32(&tmem_full[buf]));
            ait++;
        }
    } else {
        // ============ epilogue: TMEM -> registers -> Y (or split-K partial + ordered reduction by the last CTA) ============
        kf_grid_dependency_wait();  // residual / workspace / counters may still be in use by the previous kernel
        const int quad = warp & 3;
        const int row  = quad * 32 + lane;
        const int et   = (warp - kEpiWarp0) * 32 + lane;  // 0..127
        uint32_t ait   = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const Item w  = decode_item<KBR>(p, item);
            const int buf = ait % NACC;
            const int m0  = w.m0 * BN;
            const int cnt = min(BN, p.M - m0);
            const int grow = w.n0 + row;
            const int N = p.Nw[w.wi];
            void* const y = p.y[w.wi];
            const bool row_ok = grow < N;
            if (NACC == 2)
                mbar_wait_sleepy(&tmem_full[buf], (ait / NACC) & 1);  // a whole main loop away: poll politely
            else
                mbar_wait(&tmem_full[buf], (ait / NACC) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
            float* wsp = p.ws + ((size_t)w.tile * p.splits + w.z) * (size_t)(BN * BM);
            for (int c0 = 0; c0 < cnt; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tbase + (uint32_t)c0, v);
                if (p.splits > 1) {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (c0 + j < cnt) wsp[(size_t)(c0 + j) * BM + row] = __uint_as_float(v[j]);
                } else if (row_ok) {
                    float acc[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[j] = __uint_as_float(v[j]);
                    store_cols<16>(p, y, N, acc, m0 + c0, cnt - c0, grow);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);  // the accumulator buffer may be overwritten by the next item
            ait++;
            if (p.splits > 1) {
                __threadfence();
                epi_bar_sync();
                if (et == 0) {
                    const unsigned old = atomicAdd(&p.cnt[w.tile], 1u);
                    last_flag          = (old == (unsigned)(p.splits - 1));
                    if (last_flag) p.cnt[w.tile] = 0;  // self-reset for the next launch
                }
                epi_bar_sync();
                if (last_flag) {
                    __threadfence();
                    const float* base = p.ws + (size_t)w.tile * p.splits * (size_t)(BN * BM);
                    if (row_ok) {
                        // 8 tokens x 2 partials in flight per thread; the partials are added in split order (deterministic)
                        for (int j0 = 0; j0 < cnt; j0 += 8) {
                            float acc[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) acc[u] = 0.f;
                            for (int z = 0; z < p.splits; z += 2) {
                                float t0[8], t1[8];
                                const float* b0 = base + (size_t)z * (BN * BM) + (size_t)j0 * BM + row;
                                const bool two  = z + 1 < p.splits;
#pragma unroll
                                for (int u = 0; u < 8; u++) {
                                    const bool ok = j0 + u < cnt;
                                    t0[u] = ok ? __ldcg(b0 + u * BM) : 0.f;
                                    t1[u] = (ok && two) ? __ldcg(b0 + BN * BM + u * BM) : 0.f;
                                }
#pragma unroll
                                for (int u = 0; u < 8; u++) acc[u] = (acc[u] + t0[u]) + t1[u];
                            }
                            store_cols<8>(p, y, N, acc, m0 + j0, cnt - j0, grow);
                        }
                    }
                }
                epi_bar_sync();  // last_flag is rewritten by the next item
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapKey {
    const void* base;
    uint64_t inner, outer;
    uint32_t box_inner, box_outer, dt, sw, esize;
    bool operator<(const TmapKey& o) const {
        return std::tie(base, inner, outer, box_inner, box_outer, dt, sw, esize) <
               std::tie(o.base, o.inner, o.outer, o.box_inner, o.box_outer, o.dt, o.sw, o.esize);
    }
};
using TmapCache = std::map<TmapKey, CUtensorMap>;
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}
int make_map_2d(kf_ctx* ctx, CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                uint32_t box_outer, CUtensorMapSwizzle sw) {
    // a tensor map depends only on (address, geometry, box, swizzle): encode each once per context and reuse it (weights and the context's
    // activation scratch keep their addresses from token to token)
    TmapKey key = {base, inner, outer, box_inner, box_outer, (uint32_t)dt, (uint32_t)sw, (uint32_t)esize};
    TmapCache* cache = reinterpret_cast<TmapCache*>(ctx->tmap_cache);
    if (!cache) ctx->tmap_cache = cache = new TmapCache();
    auto hit = cache->find(key);
    if (hit != cache->end()) {
        *tm = hit->second;
        return KF_OK;
    }
    EncodeTiledFn fn = encode_fn();
    KF_REQUIRE(ctx, fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    KF_REQUIRE(ctx, ((uintptr_t)base & 15) == 0 && (inner * esize) % 16 == 0, "TMA needs 16-byte aligned rows");
    cuuint64_t dims[2]    = {inner, outer};
    cuuint64_t strides[1] = {inner * (uint64_t)esize};
    cuuint32_t box[2]     = {box_inner, box_outer};
    cuuint32_t estr[2]    = {1, 1};
    CUresult r = fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        ctx->last_error = b;
        return KF_ERR_CUDA;
    }
    if (cache->size() > 4096) cache->clear();  // models hold a few hundred weights; bound it anyway
    (*cache)[key] = *tm;
    return KF_OK;
}

template <int FMT, int MODE, int BN>
int launch_tc(kf_ctx* ctx, GemmParams& p, const void* const* wdata, int nw, const void* xp) {
    using F              = Fmt<FMT>;
    constexpr bool A_TMEM = FMT != TF_BF16;
    const size_t smem    = Rings<FMT, BN>::SMEM;
    auto kern            = kf_gemm_tc_kernel<FMT, MODE, BN>;
    static bool attr_set[kf_ctx::kMaxDevices] = {};  // function attributes are per device (one flag per instantiation and device)
    if (!attr_set[ctx->device]) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device] = true;
    }
    // ---- decomposition: row tiles x token tiles x split-K, walked round-robin by one CTA per SM ----
    p.n_tiles = 0;
    for (int i = 0; i <= kMaxW; i++) p.tile0[i] = 0x7fffffff;
    for (int i = 0; i < nw; i++) p.tile0[i] = p.n_tiles, p.n_tiles += (p.Nw[i] + BM - 1) / BM;
    p.m_tiles = (p.M + BN - 1) / BN;
    const int nkb = p.K / BK, nraw = (nkb + F::KBR - 1) / F::KBR;
    const int base_items = p.n_tiles * p.m_tiles;
    int splits = 1;
    if (ctx->gemv_splitk > 0) {
        splits = ctx->gemv_splitk;
    } else {
        double best = 1e30;
        for (int s = 1; s <= 16; s++) {
            if (s > nraw) break;
            const int waves  = (base_items * s + ctx->sm_count - 1) / ctx->sm_count;
            // k-blocks per item + restart of the pipeline + (split) partial write / ordered fix-up, which grows with the token tile
            const double len = (double)((nraw + s - 1) / s) * F::KBR + 2.0 + (s > 1 ? BN / 16.0 : 0.0);
            const double c   = waves * len;
            if (c < best * 0.95) best = c, splits = s;
        }
    }
    splits = std::max(1, std::min(splits, nraw));
    p.splits = splits, p.n_items = base_items * splits;
    if (splits > 1) {
        int rc = kf_ensure_gemv_ws(ctx, (size_t)p.n_items * BN * BM * sizeof(float), base_items);
        if (rc) return rc;
        p.ws = ctx->gemv_ws, p.cnt = ctx->gemv_cnt;
    }
    CUtensorMap tm_w[kMaxW], tm_x;
    int rc = KF_OK;
    for (int i = 0; i < kMaxW && !rc; i++) {
        const int j = i < nw ? i : 0;  // unused slots repeat the first weight (never dereferenced)
        if (A_TMEM)
            rc = make_map_2d(ctx, &tm_w[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, wdata[j], (uint64_t)p.K * F::BITS / 8, (uint64_t)p.Nw[j], F::RAWB, BM,
                             F::RAWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        else
            rc = make_map_2d(ctx, &tm_w[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wdata[j], (uint64_t)p.K, (uint64_t)p.Nw[j], BK, BM,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (!rc) rc = make_map_2d(ctx, &tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xp, (uint64_t)p.K, (uint64_t)p.M, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const int grid = std::min(p.n_items, ctx->sm_count);
    KF_CUDA(ctx, kf_launch_pdl(ctx, kern, dim3(grid), dim3(kThreadsTC), smem, tm_w[0], tm_w[1], tm_w[2], tm_x, p));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
template <int FMT, int MODE>
int launch_tc_bn(kf_ctx* ctx, GemmParams& p, const void* const* wdata, int nw, const void* xp) {
    if (p.M > 128) return launch_tc<FMT, MODE, 256>(ctx, p, wdata, nw, xp);
    if (p.M > 64) return launch_tc<FMT, MODE, 128>(ctx, p, wdata, nw, xp);
    if (p.M > 32) return launch_tc<FMT, MODE, 64>(ctx, p, wdata, nw, xp);
    if (p.M > 16) return launch_tc<FMT, MODE, 32>(ctx, p, wdata, nw, xp);
    return launch_tc<FMT, MODE, 16>(ctx, p, wdata, nw, xp);
}

int tc_format(const kf_tensor_desc* w, int* fmt, int* mode, int deq_fma) {
    switch (w->type) {
        case KF_T_BF16: *fmt = TF_BF16, *mode = TM_PLAIN; return KF_OK;
        case KF_T_F8E5M2: *fmt = TF_F8, *mode = TM_PLAIN; return KF_OK;
        case KF_T_Q4: *fmt = TF_Q4, *mode = deq_fma ? TM_AFFINE_FMA : w->qbias == 0 ? TM_AFFINE : TM_AFFINE_SYM; return KF_OK;
        case KF_T_Q2: *fmt = TF_Q2, *mode = deq_fma ? TM_AFFINE_FMA : w->qbias == 0 ? TM_AFFINE : TM_AFFINE_SYM; return KF_OK;
        case KF_T_SIGN: *fmt = TF_Q2, *mode = TM_SCALE; return KF_OK;
        case KF_T_BINARY: *fmt = TF_Q1, *mode = TM_SCALE; return KF_OK;
    }
    return KF_ERR_UNSUPPORTED;
}

}  // namespace
void kf_tmap_cache_destroy(kf_ctx* ctx) {
    delete reinterpret_cast<TmapCache*>(ctx->tmap_cache);
    ctx->tmap_cache = nullptr;
}

// The k order the tensor-core kernel expects for weights of w's type: identity for bf16 / f8, the extraction-friendly permutation
// inside every 32-wide slot for the packed types.  Returns x itself or the context's scratch holding the permuted copy.
int kf_tc_prepare_x(kf_ctx* ctx, const kf_tensor_desc* w, const void* x, int M, const void** xp_out) {
    int fmt, mode;
    if (tc_format(w, &fmt, &mode, 1)) return KF_ERR_UNSUPPORTED;
    if (fmt == TF_BF16 || fmt == TF_F8) {
        *xp_out = x;
        return KF_OK;
    }
    const int K         = w->cols;
    const size_t xbytes = (size_t)M * K * 2;
    int rc = kf_ensure_buf(ctx, &ctx->xperm, &ctx->xperm_bytes, xbytes);
    if (rc) return rc;
    const size_t n_slots  = (size_t)M * K / 32;
    const unsigned blocks = (unsigned)((n_slots + 255) / 256);
    if (fmt == TF_Q4)
        kf_permute_x_kernel<TF_Q4><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    else if (fmt == TF_Q2)
        kf_permute_x_kernel<TF_Q2><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    else
        kf_permute_x_kernel<TF_Q1><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    KF_LAUNCH_CHECK(ctx);
    *xp_out = ctx->xperm;
    return KF_OK;
}
// RMSNorm + kf_tc_prepare_x in one launch (packed types); bf16 / f8 weights get the plain RMSNorm into the context's xnorm scratch
int kf_tc_prepare_x_norm(kf_ctx* ctx, const kf_tensor_desc* w, const void* x, const void* norm_w, float eps, int M, const void** xp_out) {
    int fmt, mode;
    if (tc_format(w, &fmt, &mode, 1)) return KF_ERR_UNSUPPORTED;
    const int K = w->cols;
    KF_REQUIRE(ctx, K % 32 == 0, "K");
    if (fmt == TF_BF16 || fmt == TF_F8) {
        int rc = kf_ensure_buf(ctx, &ctx->xnorm, &ctx->xnorm_bytes, (size_t)M * K * 2);
        if (!rc) rc = kf_rmsnorm(ctx, ctx->xnorm, x, norm_w, M, K, eps);
        *xp_out = ctx->xnorm;
        return rc;
    }
    int rc = kf_ensure_buf(ctx, &ctx->xperm, &ctx->xperm_bytes, (size_t)M * K * 2);
    if (rc) return rc;
    if (fmt == TF_Q4)
        kf_rmsnorm_permute_kernel<TF_Q4><<<M, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, (const uint16_t*)norm_w, K, eps);
    else if (fmt == TF_Q2)
        kf_rmsnorm_permute_kernel<TF_Q2><<<M, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, (const uint16_t*)norm_w, K, eps);
    else
        kf_rmsnorm_permute_kernel<TF_Q1><<<M, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, (const uint16_t*)norm_w, K, eps);
    KF_LAUNCH_CHECK(ctx);
    *xp_out = ctx->xperm;
    return KF_OK;
}
// 0 when both weights want the same activation order (one prepared copy serves both)
int kf_tc_same_order(const kf_tensor_desc* a, const kf_tensor_desc* b) {
    int fa, fb, ma, mb;
    if (tc_format(a, &fa, &ma, 1) || tc_format(b, &fb, &mb, 1)) return 1;
    const bool pa = !(fa == TF_BF16 || fa == TF_F8), pb = !(fb == TF_BF16 || fb == TF_F8);
    if (!pa && !pb) return 0;
    return fa == fb ? 0 : 1;
}

// xp: activations as returned by kf_tc_prepare_x for weights of this type.  n <= 3 weights of the SAME storage type, K and group share
// one launch (their row tiles form one item space): Q/K/V and gate/up cost one kernel each instead of three / two.
int kf_gemm_tc_multi(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* xp, int M, int epilogue, const void* residual) {
    KF_REQUIRE(ctx, n >= 1 && n <= kMaxW && y && w && xp && M >= 1, "args");
    const int K = w[0].cols;
    int fmt, mode;
    if (tc_format(&w[0], &fmt, &mode, ctx->deq_fma)) return KF_ERR_UNSUPPORTED;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    const void* wdata[kMaxW] = {};
    p.residual = (const uint16_t*)residual, p.M = M, p.K = K;
    p.qbias = w[0].qbias, p.epilogue = epilogue;
    p.lop_mask = fmt == TF_Q4 ? 0x000F000Fu : fmt == TF_Q2 ? 0x00030003u : 0x00010001u, p.lop_magic = 0x43004300u;
    for (int i = 0; i < n; i++) {
        int f2, m2;
        KF_REQUIRE(ctx, y[i] && !tc_format(&w[i], &f2, &m2, ctx->deq_fma) && f2 == fmt && m2 == mode && w[i].cols == K && w[i].group == w[0].group &&
                            w[i].qbias == w[0].qbias,
                   "weights of one launch must share type, K, group");
        KF_REQUIRE(ctx, K % 128 == 0 && w[i].rows % 16 == 0, "K must be a multiple of 128, rows of 16");
        p.y[i] = y[i], p.Nw[i] = w[i].rows, wdata[i] = w[i].data_dev;
        if (mode != TM_PLAIN) {
            KF_REQUIRE(ctx, kf_has_gama(w[i]) && w[i].group >= 128 && (w[i].group & (w[i].group - 1)) == 0 && K % w[i].group == 0,
                       "group = 128 * 2^n dividing K");
            p.zero[i] = kf_gama_zero(w[i]), p.step[i] = kf_gama_step(w[i]);
        }
    }
    if (mode != TM_PLAIN) {
        int gs = 0;
        while ((128 << gs) < w[0].group) gs++;
        p.gshift = gs;
    }
    if (epilogue == 1) KF_REQUIRE(ctx, residual && n == 1, "residual: single weight");
#define KF_TC_CASE(F, MD) \
    if (fmt == F && mode == MD) return launch_tc_bn<F, MD>(ctx, p, wdata, n, xp);
    KF_TC_CASE(TF_Q4, TM_AFFINE_FMA)
    KF_TC_CASE(TF_Q2, TM_AFFINE_FMA)
    KF_TC_CASE(TF_Q4, TM_AFFINE)
    KF_TC_CASE(TF_Q4, TM_AFFINE_SYM)
    KF_TC_CASE(TF_Q2, TM_AFFINE)
    KF_TC_CASE(TF_Q2, TM_AFFINE_SYM)
    KF_TC_CASE(TF_Q2, TM_SCALE)
    KF_TC_CASE(TF_Q1, TM_SCALE)
    KF_TC_CASE(TF_F8, TM_PLAIN)
    KF_TC_CASE(TF_BF16, TM_PLAIN)
#undef KF_TC_CASE
    return KF_ERR_UNSUPPORTED;
}
int kf_gemm_tc(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* xp, int M, int epilogue, const void* residual) {
    void* ys[1] = {y};
    return kf_gemm_tc_multi(ctx, 1, ys, w, xp, M, epilogue, residual);
}
