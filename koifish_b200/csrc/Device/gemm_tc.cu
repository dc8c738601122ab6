// gemm_tc.cu -- dequant-fused tensor-core GEMM for prefill / large batches (M > 64 tokens) on the 5th-generation tensor cores:
//   Y[M][N] = X[M][K] . deq(W[N][K])^T       tcgen05.mma (kind::f16, bf16 x bf16 -> fp32), accumulators in TMEM
//
// Replaces, for nToken >= 65, the same reference pair as gemv.cu: GTensor::GetDataX (whole-matrix dequant to a bf16 scratch,
// src/Device/CUDA/kernel/quantizer.cu:249-392) + CU_mm_blasLt (src/Device/CUDA/kernel/gemm.cu:93-214) behind SLP::Forw
// (src/Device/CUDA/NeuronFuse.cu:305-381).  The packed weights are read once per 128/256-token panel and expanded on chip.
//
// Tiling: one CTA computes a [128 weight rows] x [BN tokens] tile of Y^T.  The weights are the A operand (UMMA_M = 128 rows), the
// activations the B operand (UMMA_N = BN tokens), both K-major in shared memory in the canonical SWIZZLE_128B layout, BK = 64:
//   * warps 0-3 (128 threads, one weight row each): load the row's packed words of the k-block, expand them to the reference's
//     bit-exact bf16 weights (same two-rounding arithmetic as gemv.cu) and store the 128-byte row into the swizzled A tile; copy the
//     activation rows into the B tile with cp.async.  The k-order inside each 32-wide slot is the extraction-friendly permutation of
//     gemv.cu; the activations are permuted identically beforehand by kf_permute_x_kernel, so the contraction is unchanged;
//   * warp 4: one elected thread issues tcgen05.mma for the stage (4 x K16), tcgen05.commit releases the stage / signals the end;
//   * warps 0-3 then read the accumulator from TMEM (tcgen05.ld 32x32b: lane = weight row, column = token) and write Y.
// Pipeline: STAGES-deep ring of {A, B} tiles guarded by full/empty mbarriers.  Every wait is bounded and traps instead of hanging.
#include <string.h>

#include <algorithm>

#include "kf_common.cuh"

namespace {

enum { TF_BF16 = 0, TF_F8 = 1, TF_Q4 = 2, TF_Q2 = 3, TF_Q1 = 4 };
enum { TM_PLAIN = 0, TM_AFFINE = 1, TM_AFFINE_SYM = 2, TM_SCALE = 3 };

constexpr int BM = 128;  // weight rows per CTA (UMMA M)
constexpr int BK = 64;   // k per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreadsTC       = kProducerThreads + 32;

struct GemmParams {
    const uint8_t* data;
    const uint16_t* zero;
    const uint16_t* step;
    const uint16_t* xp;        // activations, permuted inside every 32-wide k slot: [M][K]
    void* y;                   // bf16 [M][N] (or float when epilogue == 4)
    const uint16_t* residual;  // bf16 [M][N] or nullptr
    int M, N, K;
    int qbias, gshift, epilogue;
    uint32_t lop_mask, lop_magic;
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

// ---- PTX helpers ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); it++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
    const uint32_t d = smem_u32(dst);
    const int sz     = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

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
    } else if constexpr (FMT == TF_F8) {  // 8 registers, natural order
#pragma unroll
        for (int r = 0; r < 8; r++) o[2 * r] = tf8_pair(w[r], 0x1404u), o[2 * r + 1] = tf8_pair(w[r], 0x3424u);
    } else {  // bf16: 16 registers, natural order
#pragma unroll
        for (int r = 0; r < 16; r++) o[r] = w[r];
    }
}

template <int FMT, int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(kThreadsTC, 1) kf_gemm_tc_kernel(const GemmParams p) {
    constexpr int BITS    = FMT == TF_BF16 ? 16 : FMT == TF_F8 ? 8 : FMT == TF_Q4 ? 4 : FMT == TF_Q2 ? 2 : 1;
    constexpr int SLOTB   = 32 * BITS / 8;  // packed bytes of one 32-weight slot
    constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned tiles (the 128-byte swizzle is a function of address bits [4,10))
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BM, m0 = blockIdx.y * BN;
    const int nkb = p.K / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], kProducerThreads), mbar_init(&empty_bar[s], 1);
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {  // TMEM: BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp < 4) {
        // =================================== producers: one weight row / one token row per thread ===================================
        const int r           = tid;  // row inside the tile
        const int grow        = min(n0 + r, p.N - 1);
        const size_t row_bytes = (size_t)p.K * BITS / 8;
        const uint8_t* wrow   = p.data + (size_t)grow * row_bytes;
        const int gpr         = (p.K >> 7) >> p.gshift;
        const uint16_t* zrow  = MODE == TM_PLAIN ? nullptr : p.zero + (size_t)grow * gpr;
        const uint16_t* srow  = MODE == TM_PLAIN ? nullptr : p.step + (size_t)grow * gpr;
        const uint32_t bias2  = pack_bf16x2((float)(128 + p.qbias), (float)(128 + p.qbias));
        const int mrow        = m0 + r;  // token row this thread copies (BN == 128: one each; BN == 256: two each)
        const uint32_t sw     = (uint32_t)(r & 7);
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % STAGES;
            mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
            uint8_t* a_tile = tiles + (size_t)s * (A_BYTES + B_BYTES);
            uint8_t* b_tile = a_tile + A_BYTES;
            // ---- B: activation rows (already permuted), 128 bytes each, async ----
#pragma unroll
            for (int rep = 0; rep < BN / 128; rep++) {
                const int mr      = r + rep * 128;
                const int gm      = mrow + rep * 128;
                const bool valid  = gm < p.M;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.xp + (size_t)(valid ? gm : 0) * p.K + (size_t)kb * BK);
                uint8_t* dst       = b_tile + (mr >> 3) * 1024 + (mr & 7) * 128;
#pragma unroll
                for (int c = 0; c < 8; c++) cp_async16_zfill(dst + ((c ^ (mr & 7)) << 4), src + c * 16, valid);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // ---- A: this row's two 32-weight slots of the k-block -> 64 bf16 = one swizzled 128-byte row ----
            uint8_t* arow = a_tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
            for (int slot = 0; slot < 2; slot++) {
                const int kslot = kb * 2 + slot;  // 32-wide slot index along K
                uint32_t w[SLOTB / 4 > 0 ? SLOTB / 4 : 1];
                int off = kslot * SLOTB;
                if (FMT == TF_Q4) {
                    off = (kslot >> 2) * 64 + (kslot & 3) * 16;  // word t of the group
                } else if (FMT == TF_Q2) {
                    const int g4 = kslot & 3;                    // word.high holds the first 32 codes (PackedQ.hpp:185-198)
                    off = (kslot >> 2) * 32 + 16 * (g4 >> 1) + 8 * (1 - (g4 & 1));
                } else if (FMT == TF_Q1) {
                    off = (kslot >> 2) * 16 + 12 - 4 * (kslot & 3);  // high.hi32 holds codes 0..31 (PackedQ.hpp:200-211)
                }
                if constexpr (SLOTB >= 16) {
#pragma unroll
                    for (int j = 0; j < SLOTB / 16; j++) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + off) + j);
                        w[4 * j] = v.x, w[4 * j + 1] = v.y, w[4 * j + 2] = v.z, w[4 * j + 3] = v.w;
                    }
                } else if constexpr (SLOTB == 8) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(wrow + off));
                    w[0] = v.x, w[1] = v.y;
                } else {
                    w[0] = __ldg(reinterpret_cast<const uint32_t*>(wrow + off));
                }
                uint32_t step2 = 0, zero2 = 0, nb2 = 0;
                if (MODE != TM_PLAIN) {
                    const int gi      = (kslot >> 2) >> p.gshift;
                    const uint32_t zz = __ldg(zrow + gi), ss = __ldg(srow + gi);
                    step2 = ss | (ss << 16), zero2 = zz | (zz << 16);
                    if (MODE == TM_AFFINE) nb2 = bf162_as_u32(__hmul2_rn(u32_as_bf162(step2), u32_as_bf162(0xC300C300u)));
                }
                uint32_t o[16];
                expand_slot<FMT, MODE>(o, w, step2, zero2, nb2, bias2, p.lop_mask, p.lop_magic);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t chunk = (uint32_t)(slot * 4 + c) ^ sw;
                    *reinterpret_cast<uint4*>(arow + (chunk << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            fence_proxy_async();  // generic-proxy writes (st.shared, cp.async) -> visible to the tensor core's async proxy
            mbar_arrive(&full_bar[s]);
        }
        // =================================== epilogue: TMEM -> registers -> Y ===================================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int row       = n0 + warp * 32 + lane;  // TMEM lane == weight row; warp w owns lanes 32w .. 32w+31
        const bool row_ok   = row < p.N;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
                "%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int m = m0 + c0 + j;
                    if (m >= p.M) break;
                    const float acc = __uint_as_float(v[j]);
                    const size_t idx = (size_t)m * p.N + row;
                    if (p.epilogue == 4) {
                        reinterpret_cast<float*>(p.y)[idx] = acc;
                    } else {
                        uint16_t b = f32_to_bf16_bits(acc);  // the reference's GEMM output is bf16 (gemm.cu:124-126)
                        if (p.epilogue == 1) b = f32_to_bf16_bits(bf16_bits_to_f32(p.residual[idx]) + bf16_bits_to_f32(b));
                        reinterpret_cast<uint16_t*>(p.y)[idx] = b;
                    }
                }
            }
        }
        tc_fence_before();
    } else if (lane == 0) {
        // =================================== MMA issuer: a single thread ===================================
        constexpr uint32_t idesc = make_idesc_bf16(BN);
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % STAGES;
            mbar_wait(&full_bar[s], (kb / STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(tiles + (size_t)s * (A_BYTES + B_BYTES));
            const uint64_t adesc = make_desc_sw128(a_addr), bdesc = make_desc_sw128(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; k++)  // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in 16-byte units
                umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);  // arrives when the MMAs above have finished reading the stage
        }
        umma_commit(&tmem_full_bar);
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)BN) : "memory");
    }
}

template <int FMT, int MODE, int BN>
int launch_tc(kf_ctx* ctx, const GemmParams& p) {
    constexpr int STAGES = BN == 256 ? 4 : 6;
    const size_t smem    = (size_t)STAGES * (BM * 128 + BN * 128) + 1024;
    auto kern            = kf_gemm_tc_kernel<FMT, MODE, BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid((p.N + BM - 1) / BM, (p.M + BN - 1) / BN);
    kern<<<grid, kThreadsTC, smem, ctx->stream>>>(p);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
template <int FMT, int MODE>
int launch_tc_bn(kf_ctx* ctx, const GemmParams& p) {
    return p.M > 128 ? launch_tc<FMT, MODE, 256>(ctx, p) : launch_tc<FMT, MODE, 128>(ctx, p);
}

}  // namespace

// xp_scratch: device buffer of M*K bf16 for the permuted activations (owned by the caller / context)
int kf_gemm_tc(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    KF_REQUIRE(ctx, y && w && x && M >= 1, "args");
    const int K = w->cols, N = w->rows;
    KF_REQUIRE(ctx, K % 128 == 0 && N % 16 == 0, "K must be a multiple of 128, rows of 16");
    int fmt, mode;
    switch (w->type) {
        case KF_T_BF16: fmt = TF_BF16, mode = TM_PLAIN; break;
        case KF_T_F8E5M2: fmt = TF_F8, mode = TM_PLAIN; break;
        case KF_T_Q4: fmt = TF_Q4, mode = w->qbias == 0 ? TM_AFFINE : TM_AFFINE_SYM; break;
        case KF_T_Q2: fmt = TF_Q2, mode = w->qbias == 0 ? TM_AFFINE : TM_AFFINE_SYM; break;
        case KF_T_SIGN: fmt = TF_Q2, mode = TM_SCALE; break;
        case KF_T_BINARY: fmt = TF_Q1, mode = TM_SCALE; break;
        default: return KF_ERR_UNSUPPORTED;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.data = (const uint8_t*)w->data_dev, p.y = y, p.residual = (const uint16_t*)residual, p.M = M, p.N = N, p.K = K;
    p.qbias = w->qbias, p.epilogue = epilogue;
    p.lop_mask = fmt == TF_Q4 ? 0x000F000Fu : fmt == TF_Q2 ? 0x00030003u : 0x00010001u, p.lop_magic = 0x43004300u;
    if (mode != TM_PLAIN) {
        KF_REQUIRE(ctx, kf_has_gama(*w) && w->group >= 128 && (w->group & (w->group - 1)) == 0 && K % w->group == 0, "group = 128 * 2^n dividing K");
        p.zero = kf_gama_zero(*w), p.step = kf_gama_step(*w);
        int gs = 0;
        while ((128 << gs) < w->group) gs++;
        p.gshift = gs;
    }
    if (epilogue == 1) KF_REQUIRE(ctx, residual, "residual");
    // permute the activations once per GEMM (shared by every row tile)
    const size_t xbytes = (size_t)M * K * 2;
    int rc = kf_ensure_buf(ctx, &ctx->xperm, &ctx->xperm_bytes, xbytes);
    if (rc) return rc;
    p.xp = (const uint16_t*)ctx->xperm;
    const size_t n_slots = (size_t)M * K / 32;
    const unsigned blocks = (unsigned)((n_slots + 255) / 256);
    if (fmt == TF_Q4)
        kf_permute_x_kernel<TF_Q4><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    else if (fmt == TF_Q2)
        kf_permute_x_kernel<TF_Q2><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    else if (fmt == TF_Q1)
        kf_permute_x_kernel<TF_Q1><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)ctx->xperm, (const uint16_t*)x, n_slots);
    else
        p.xp = (const uint16_t*)x;  // byte / bf16 streams keep the natural order
    if (p.xp != (const uint16_t*)x) KF_LAUNCH_CHECK(ctx);
#define KF_TC_CASE(F, MD) \
    if (fmt == F && mode == MD) return launch_tc_bn<F, MD>(ctx, p);
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
