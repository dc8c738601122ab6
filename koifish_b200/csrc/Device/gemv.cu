// gemv.cu -- dequant-fused GEMV / skinny GEMM for decode (M <= 64 tokens):  y[M][N] = x[M][K] . deq(W[N][K])^T
//
// Replaces the reference pair  GTensor::GetDataX (whole-matrix dequant to a bf16 scratch, src/Device/CUDA/kernel/quantizer.cu:
// 249-392, T.cu:245-294)  +  CU_mm_blasLt (cuBLASLt bf16 GEMM, src/Device/CUDA/kernel/gemm.cu:93-214)  as called from
// SLP::Forw (src/Device/CUDA/NeuronFuse.cu:305-381), plus the launches around it: the preceding CU_rms_infer (layernorm.cuh:
// 801-859) can be folded into the activation staging, the following CU_swiglu_v0 / CU_add3 into the epilogue.
//
// Design (HBM-bound; roofline = bytes of packed weights + gama):
//   * every packed byte is read from HBM exactly once.  Each thread copies the 16/8/4-byte words it will consume into its own slot
//     of a shared-memory ring with cp.async (no register staging, DEPTH k-steps in flight per thread, no block barrier: a thread
//     only ever reads what it copied itself).  The first DEPTH copies are issued before the prologue, so the weight stream runs
//     while the activations are being normalised / staged;
//   * the codes are expanded in registers to the *bit-exact* bf16 weights the reference's dequant kernel produces
//     (p = RN(step*k), w = RN(p - zero), both bf16) and fed directly as the A fragments of mma.sync.m16n8k16 (bf16 in, fp32
//     accumulate -- the accumulation type cuBLASLt uses in the reference).  Tensor-core *throughput* is irrelevant here (the op
//     is bandwidth bound, the MMA pipe runs at < 20 %); the MMA is used because it takes bf16 pairs without an unpack-to-fp32
//     and does the 16x8x16 FMAs in one issue slot -- the kernel is issue-slot limited, not DRAM limited, once the loads are
//     deep enough (profiles/r01_ncu_gemv_q4_v1.txt);
//   * the k-order inside a 128-wide group is permuted to make code extraction cheap (nibbles 16 bits apart form one bf16x2
//     register); the activations are staged in shared memory in the same permuted order, so no weight is ever shuffled;
//   * per-group zero/step are staged once per CTA in shared memory with wide batched loads;
//   * a warp owns RT row tiles of 16 rows that share every activation fragment;
//   * split-K across CTAs with a deterministic (fixed-order) last-CTA reduction; epilogues fuse the residual add or SwiGLU.
#include <string.h>

#include <algorithm>

#include "kf_common.cuh"
#include "kf_tp.cuh"

namespace {

enum { FMT_BF16 = 0, FMT_F8 = 1, FMT_Q4 = 2, FMT_Q2 = 3, FMT_Q1 = 4 };
enum { MODE_PLAIN = 0, MODE_AFFINE = 1, MODE_AFFINE_SYM = 2, MODE_SCALE = 3, MODE_FACTOR = 4, MODE_AFFINE_FMA = 5, MODE_FAST = 6 };
// MODE_FAST (4 / 2 / 1-bit, the DEFAULT for decode: ctx knob gemv_exact = 0): the codes go to the tensor cores as fp16 SUBNORMALS.  A code
// field that lies inside the low 10 bits of a 16-bit half IS the fp16 number c * 2^b * 2^-24 (exponent field 0: no implicit one, no offset to
// cancel), so ONE LOP3 (`word & mask`, mask an immediate) turns two packed codes into an MMA operand pair; the fields above bit 9 come down
// with one byte permute (>> 8) per register.  mma.sync handles fp16 subnormals exactly (tools/ubench/hmma_subnormal.cu).  The activation that
// meets the field at bit b is staged as fp16 x * 2^-b (exact), with a power-of-two scale per 128-k group that puts the group's largest
// magnitude in [2^13, 2^14).  Per weight that is 0.56 .. 0.62 integer-pipe instructions instead of 3 (shift + LOP3 + bf16 subtract of the
// bf16 forms; a funnel shift alone costs ~5 clk per warp instruction and scheduler on B200, tools/ubench/deq_rate2.cu: the bit-exact forms
// need 432 clk per [128 x 128] tile, MORE than the tile's bytes take at the HBM rate).  The group's step / zero / code bias are applied to
// the fp32 group sums, y += step * (sum c x - qbias Sx) - zero Sx: the reference's affine dequant without its per-weight bf16 rounding (for
// the yyang ternary / binary types, whose weights step * k are exact in bf16, that is the SAME arithmetic as the reference's).
// MODE_AFFINE / MODE_AFFINE_SYM: the reference's dequant with TWO bf16 roundings (ctx knob deq_fma = 0) ; MODE_AFFINE_FMA: with ONE
// (fma.rn.bf16, the default: what the reference's kernel computes when built for sm_90+, see kf_common.cuh deq_fma)
// MODE_FACTOR (opt-in, ctx knob gemv_exact = 0): A = 128 + code, un-dequantised; per group y += step*(acc_g - (128+qbias)*Sx) - zero*Sx with
// Sx = sum of the group's activations.  Mathematically the same affine map, but WITHOUT the reference's two bf16 roundings of the
// weights (result differs from the exact modes by about one bf16 ulp of y); it exists to measure what the roundings cost.
enum { EPI_NONE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_F32 = 4, EPI_TP = 5 /* fused tensor-parallel exchange, kf_tp.cuh */ };

constexpr int kThreads = 256;
constexpr int KSTEP    = 128;  // k per step for every format (= one quantisation group of the packed formats)
constexpr int UNITS    = 4;    // 8-element activation units per thread slot and step
constexpr int KT       = 32;   // k per thread slot

// CPB: bytes per cp.async ; NCH: copies per (row, k-step) and thread ; D1 / D2: ring depth for 16 / 32 rows per warp
template <int FMT> struct Fmt;
// cp.async ring depths (k-steps in flight per thread).  Measured inside the decode step (profiles/r01_ring_depth.txt): a SHALLOW ring wins --
// 3 stages for 4-bit give 155 tok/s on Qwen3-32B against 145 with 5 and 114 with 8: the shared memory a CTA does not take lets the next
// kernel's CTAs become resident (and start their own weight stream) before this kernel has drained.
#ifndef KF_GEMV_D1_Q4
#define KF_GEMV_D1_Q4 3
#endif
template <> struct Fmt<FMT_Q4>   { static constexpr int BITS = 4,  CPB = 16, NCH = 1, D1 = KF_GEMV_D1_Q4,  D2 = 3; };
#ifndef KF_GEMV_D1_Q2
#define KF_GEMV_D1_Q2 4
#endif
#ifndef KF_GEMV_D1_Q1
#define KF_GEMV_D1_Q1 6
#endif
template <> struct Fmt<FMT_Q2>   { static constexpr int BITS = 2,  CPB = 8,  NCH = 1, D1 = KF_GEMV_D1_Q2,  D2 = 5; };
template <> struct Fmt<FMT_Q1>   { static constexpr int BITS = 1,  CPB = 4,  NCH = 1, D1 = KF_GEMV_D1_Q1, D2 = 8; };
template <> struct Fmt<FMT_F8>   { static constexpr int BITS = 8,  CPB = 16, NCH = 2, D1 = 3,  D2 = 2; };
template <> struct Fmt<FMT_BF16> { static constexpr int BITS = 16, CPB = 16, NCH = 4, D1 = 2,  D2 = 2; };

__host__ __device__ constexpr int fmt_tb(int bits) { return KT * bits / 8; }  // bytes per thread, row and k-step

struct GemvSeg {
    const uint8_t* data;
    const uint16_t* zero;
    const uint16_t* step;
    uint16_t* y;
    int rows;
    int rb0;  // first row block of this segment
};
struct GemvParams {
    GemvSeg seg[3];
    int nseg;
    const uint16_t* x;
    const uint16_t* residual;
    const uint16_t* norm_w;  // optional fused RMSNorm of x (weights [K]); nullptr = x is used as is
    float norm_eps;
    int M, K;
    int steps_total;  // K / KSTEP
    int S;            // k-splits
    int cluster;      // 1: the S CTAs of a row block form a cluster and merge through distributed shared memory (M = 1)
    int red_off;      // byte offset of the leader's [S][ROWS] fp32 merge area in dynamic shared memory
    int nsteps_max;   // ceil(steps_total / S): sizes the shared-memory regions
    int total_rb;
    int qbias;
    int ring_off;  // byte offset of the weight ring inside dynamic shared memory (16-byte aligned)
    int sx_off;    // byte offset of the activation group sums (MODE_FACTOR)
    int gshift;    // log2(group / 128): gama index = row*(K/group) + (step >> gshift)
    int epilogue;
    uint32_t lop_mask, lop_magic;  // code-field mask (0x000F000F / 0x00030003 / 0x00010001) and bf16x2 128.0 (0x43004300)
    // activations pre-staged in GLOBAL memory by kf_gemv_xprep_kernel (XG variants: 4..8-token decode): same layout as the shared-memory
    // staging area, all k-steps of K
    const uint4* xg;
    const float4* sxg;
    uint4* xg_out;    // xprep only
    float4* sxg_out;  // xprep only
    int prep_steps;   // xprep only: k-steps per CTA
    float* ws;
    unsigned* cnt;
    // fused tensor-parallel exchange (kf_tp.cuh), epilogue EPI_TP: y = residual + sum over the ranks of this matmul, as exchange #tp_out
    KfTpView tp;
    int tp_out;
};

// ---- permuted k-order of the activations inside one thread slot (see header) -------------------------------------------------
template <int FMT>
__host__ __device__ constexpr int xperm(int o) {
    const int u = o >> 3, e = o & 7;
    if (FMT == FMT_Q4) return 8 * u + ((e & 1) ? 3 : 7) - (e >> 1);                         // {7,3,6,2,5,1,4,0}
    if (FMT == FMT_Q2) return 16 * (u >> 1) + ((e & 1) ? 7 : 15) - (e >> 1) - 4 * (u & 1);  // {15,7,14,6,13,5,12,4} / -4
    if (FMT == FMT_Q1) return ((e & 1) ? 15 : 31) - (e >> 1) - 4 * u;                       // {31,15,30,14,29,13,28,12} - 4u
    return o;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// MODE_FAST: bit position (inside each 16-bit half of a packed register) of the code that feeds MMA slot j = 2h + {0: a[0]/a[1], 1: a[2]/a[3]}
// of unit u, and the exponent of the fp16 subnormal it becomes (fields whose top bit lies above bit 9 are used from the register >> 8)
template <int BITS>
__host__ __device__ constexpr int fast_bit(int u, int j) {
    return BITS == 4 ? 8 * (j >> 1) + 4 * (j & 1) : BITS == 2 ? 8 * (u & 1) + 4 * (j >> 1) + 2 * (j & 1) : 4 * u + 2 * (j >> 1) + (j & 1);
}
template <int BITS>
__host__ __device__ constexpr int fast_exp(int u, int j) {
    return fast_bit<BITS>(u, j) + BITS <= 10 ? fast_bit<BITS>(u, j) : fast_bit<BITS>(u, j) - 8;
}
// the pair of codes at bit b of both halves as fp16x2 subnormals: c * 2^fast_exp * 2^-24
template <int BITS>
__device__ __forceinline__ uint32_t fast_pair(uint32_t r, uint32_t r8, int u, int j) {
    const int b         = fast_bit<BITS>(u, j);
    const uint32_t mask = ((1u << BITS) - 1u) * 0x00010001u;
    return b + BITS <= 10 ? (r & (mask << b)) : (r8 & (mask << (b - 8)));
}

// asynchronous global -> shared copy (LDGSTS): no register staging, so many k-steps can be in flight per thread
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// (a & b) | c in ONE LOP3: mask and magic come from kernel parameters so that ptxas keeps them as register / constant-bank
// operands instead of splitting the op into two immediate-form LOP3s.
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// Expand one pair of codes (16 bits apart in `reg` after the shift) to the bf16x2 weights.  The _rn intrinsics are never
// contracted by ptxas: the two roundings of the reference build (bf16 multiply, then bf16 subtract) are preserved.
template <int MODE>
__device__ __forceinline__ uint32_t deq_pair(uint32_t reg, int shift, uint32_t step2, uint32_t zero2, uint32_t nbias2, uint32_t bias2,
                                             uint32_t mask, uint32_t magic) {
    uint32_t v = and_or(reg >> shift, mask, magic);  // bf16x2 {128 + c_lo, 128 + c_hi}, exact
    if (MODE == MODE_FACTOR) return v;
    if (MODE == MODE_AFFINE) {
        // qbias == 0:  RN(step*(v-128)) == fma(v, step, -128*step) (single rounding of the exact product step*c) ; then RN(p - zero)
        __nv_bfloat162 p = __hfma2(u32_as_bf162(v), u32_as_bf162(step2), u32_as_bf162(nbias2));
        return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
    } else if (MODE == MODE_AFFINE_FMA) {
        __nv_bfloat162 k = __hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2));  // exact small integer
        return bf162_as_u32(__hfma2(k, u32_as_bf162(step2), u32_as_bf162(zero2)));  // zero2 holds -zero: RN(step*k - zero), one rounding
    } else if (MODE == MODE_AFFINE_SYM) {
        __nv_bfloat162 k = __hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2));  // exact small integer
        __nv_bfloat162 p = __hmul2_rn(u32_as_bf162(step2), k);
        return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
    } else {  // MODE_SCALE: A = code - qbias ; the group step is applied to the fp32 group sum
        return bf162_as_u32(__hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2)));
    }
}
// E5M2-by-truncation bytes -> bf16x2 (exact): fp16 bits (b<<8) re-biased into bf16 via a 2^112 multiply
__device__ __forceinline__ uint32_t f8_pair(uint32_t reg, uint32_t sel) {
    uint32_t h2 = __byte_perm(reg, 0u, sel);  // {b_hi<<8 : b_lo<<8} as fp16x2
    uint32_t t  = ((h2 >> 3) & 0x0FE00FE0u) | (h2 & 0x80008000u);
    return bf162_as_u32(__hmul2_rn(u32_as_bf162(t), u32_as_bf162(0x77807780u)));  // * 2^112
}

// A fragments of mma #(2u+h) of the current k-step for one row tile.  wa / wb: the thread's TB bytes of rows g / g+8 as 32-bit
// registers in MEMORY order (the 128-bit word formats keep the first codes in the LAST register: PackedQ.hpp:28-31).
template <int FMT, int MODE, int NR>
__device__ __forceinline__ void build_a(uint32_t (&a)[4], const uint32_t (&wa)[NR], const uint32_t (&wb)[NR], int u, int h, const uint32_t (&gm)[6],
                                        uint32_t bias2, uint32_t mask, uint32_t magic) {
    // gm = {step2a, zero2a, nb2a, step2b, zero2b, nb2b}
    if constexpr (MODE == MODE_FAST) {
        constexpr int BITS = Fmt<FMT>::BITS;
        const int ri       = BITS == 4 ? 3 - u : BITS == 2 ? 1 - (u >> 1) : 0;
        const uint32_t ra = wa[ri], rb = wb[ri];
        const uint32_t ra8 = __byte_perm(ra, 0u, 0x4321), rb8 = __byte_perm(rb, 0u, 0x4321);  // >> 8 as a byte permute (shared by the units)
        a[0] = fast_pair<BITS>(ra, ra8, u, 2 * h), a[1] = fast_pair<BITS>(rb, rb8, u, 2 * h);
        a[2] = fast_pair<BITS>(ra, ra8, u, 2 * h + 1), a[3] = fast_pair<BITS>(rb, rb8, u, 2 * h + 1);
    } else if constexpr (FMT == FMT_Q4) {
        const uint32_t ra = wa[3 - u], rb = wb[3 - u];
        a[0] = deq_pair<MODE>(ra, 8 * h, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[1] = deq_pair<MODE>(rb, 8 * h, gm[3], gm[4], gm[5], bias2, mask, magic);
        a[2] = deq_pair<MODE>(ra, 8 * h + 4, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[3] = deq_pair<MODE>(rb, 8 * h + 4, gm[3], gm[4], gm[5], bias2, mask, magic);
    } else if constexpr (FMT == FMT_Q2) {
        const uint32_t ra = wa[1 - (u >> 1)], rb = wb[1 - (u >> 1)];
        const int m4 = 4 * (2 * (u & 1) + h);
        a[0] = deq_pair<MODE>(ra, m4, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[1] = deq_pair<MODE>(rb, m4, gm[3], gm[4], gm[5], bias2, mask, magic);
        a[2] = deq_pair<MODE>(ra, m4 + 2, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[3] = deq_pair<MODE>(rb, m4 + 2, gm[3], gm[4], gm[5], bias2, mask, magic);
    } else if constexpr (FMT == FMT_Q1) {
        const uint32_t ra = wa[0], rb = wb[0];
        const int m2 = 2 * (2 * u + h);
        a[0] = deq_pair<MODE>(ra, m2, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[1] = deq_pair<MODE>(rb, m2, gm[3], gm[4], gm[5], bias2, mask, magic);
        a[2] = deq_pair<MODE>(ra, m2 + 1, gm[0], gm[1], gm[2], bias2, mask, magic);
        a[3] = deq_pair<MODE>(rb, m2 + 1, gm[3], gm[4], gm[5], bias2, mask, magic);
    } else if constexpr (FMT == FMT_F8) {  // 32 bytes per thread and row: register 2u+h holds k = 8u+4h .. +3
        const uint32_t ra = wa[2 * u + h], rb = wb[2 * u + h];
        a[0] = f8_pair(ra, 0x1404u), a[1] = f8_pair(rb, 0x1404u);
        a[2] = f8_pair(ra, 0x3424u), a[3] = f8_pair(rb, 0x3424u);
    } else {  // FMT_BF16: 64 bytes per thread and row: registers 4u+2h, 4u+2h+1 hold k = 8u+4h .. +3
        a[0] = wa[4 * u + 2 * h], a[1] = wb[4 * u + 2 * h];
        a[2] = wa[4 * u + 2 * h + 1], a[3] = wb[4 * u + 2 * h + 1];
    }
}

// ---- activation staging, shared by the matmul kernel (destination: its shared memory, one k-slice) and by kf_gemv_xprep_kernel (destination:
//      global memory, all of K, once per launch sequence).  dstx[((s * UNITS + u) * MX + m) * 4 + slot] = the 8 activations unit u / thread slot
//      `slot` of k-step s_begin + s needs, in fragment order; dstsx[s * MX + m] = the group sums of MODE_FAST.  256 threads.
template <int FMT, int MODE, int MX>
__device__ __forceinline__ void gemv_stage_x(uint4* __restrict__ dstx, float4* __restrict__ dstsx, const GemvParams& p, int s_begin, int nsteps,
                                             float* s_red, float* s_scale) {
    const int tid = threadIdx.x, lane = tid & 31;
    // ---- optional fused RMSNorm: the same arithmetic, in the same order, as kf_rmsnorm_kernel (ops.cu) --------------------------
    if (p.norm_w) {
        for (int m = 0; m < p.M; m++) {
            const uint16_t* xr = p.x + (size_t)m * p.K;
            float ss = 0.f;
            for (int i = tid * 8; i < p.K; i += kThreads * 8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(xr + i));
                const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float a = bf16lo(q[j]), b = bf16hi(q[j]);
                    ss = fmaf(a, a, ss), ss = fmaf(b, b, ss);
                }
            }
            ss = block_sum(ss, s_red);
            if (tid == 0) s_scale[m] = 1.0f / sqrtf(fmaf(ss, 1.0f / (float)p.K, p.norm_eps));
        }
        __syncthreads();
    }

    // ---- stage the activations of this k-slice in shared memory, permuted to the fragment order --------------------------------
    {
        const int items = MX * nsteps * 4;
        for (int it = tid; it < items; it += kThreads) {
            const int tt = it & 3, m = (it >> 2) % MX, s = it / (4 * MX);
            uint32_t src[KT / 2];
            if (m < p.M) {
                const size_t k0 = (size_t)(s_begin + s) * KSTEP + tt * KT;
                const uint4* gp = reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + k0);
#pragma unroll
                for (int i = 0; i < KT / 8; i++) {
                    uint4 v = __ldg(gp + i);
                    src[4 * i + 0] = v.x, src[4 * i + 1] = v.y, src[4 * i + 2] = v.z, src[4 * i + 3] = v.w;
                }
                if (p.norm_w) {  // (x * s) * w, rounded to bf16 like the stand-alone kernel's output
                    const float sc  = s_scale[m];
                    const uint4* wp = reinterpret_cast<const uint4*>(p.norm_w + k0);
#pragma unroll
                    for (int i = 0; i < KT / 8; i++) {
                        const uint4 wv = __ldg(wp + i);
                        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t xv = src[4 * i + j];
                            src[4 * i + j]    = pack_bf16x2((bf16lo(xv) * sc) * bf16lo(ww[j]), (bf16hi(xv) * sc) * bf16hi(ww[j]));
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < KT / 2; i++) src[i] = 0u;
            }
            float fscl = 1.0f;  // MODE_FAST: the group's power-of-two scale
            if (MODE == MODE_FAST) {
                // fp16 staging with a power-of-two scale per 128-k group (the quad's 4 x 32 values; the 4 lanes of a quad are always
                // active together): the group's largest magnitude lands in [2^13, 2^14).  sxs[s][m] = {0, Sx, 2^24 / scale} with
                // Sx = 2^-24 x the group sum of the scaled activations (the MMA's sums carry the 2^-24 of the subnormal codes)
                const unsigned qm = 0xFu << (lane & ~3);
                uint32_t amax = 0;
#pragma unroll
                for (int i = 0; i < KT / 2; i++) amax = max(amax, max(src[i] & 0x7fffu, (src[i] >> 16) & 0x7fffu));
                amax = max(amax, __shfl_xor_sync(qm, amax, 1));
                amax = max(amax, __shfl_xor_sync(qm, amax, 2));
                const int shift = amax == 0 ? 0 : max(-100, min(100, 140 - (int)(amax >> 7)));  // 140 = 127 + 13
                fscl = __uint_as_float((uint32_t)(127 + shift) << 23);
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < KT / 2; i++) sum += bf16lo(src[i]) * fscl, sum += bf16hi(src[i]) * fscl;
                sum += __shfl_xor_sync(qm, sum, 1), sum += __shfl_xor_sync(qm, sum, 2);  // fixed order: bit-reproducible
                sum *= 5.9604644775390625e-8f;                                            // 2^-24
                if (tt == 0)
                    dstsx[s * MX + m] =
                        make_float4(0.f, sum, __uint_as_float((uint32_t)(127 + 24 - shift) << 23), 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNITS; u++) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int e0 = xperm<FMT>(u * 8 + 2 * j), e1 = xperm<FMT>(u * 8 + 2 * j + 1);
                    uint32_t lo = (src[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu;
                    uint32_t hi = (src[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu;
                    if (MODE == MODE_FAST) {  // fp16(x * scale * 2^-b): b = the bit position of the code field this element meets
                        const float f = fscl * __uint_as_float((uint32_t)(127 - fast_exp<Fmt<FMT>::BITS>(u, j)) << 23);
                        lo = __half_as_ushort(__float2half_rn(bf16_bits_to_f32(lo) * f));
                        hi = __half_as_ushort(__float2half_rn(bf16_bits_to_f32(hi) * f));
                    }
                    o[j] = lo | (hi << 16);
                }
                // destination (unit, thread slot): packed formats keep the 32-k slot of thread tt; the byte / bf16 streams interleave
                // 16-byte chunks across the quad (chunk ch of thread t covers bytes (4*ch + t)*16 of the k-step)
                int du = u, dt = tt;
                if (FMT == FMT_BF16) {
                    du = tt, dt = u;  // natural 8-k block B = 4*tt + u  ->  unit B >> 2, thread B & 3
                } else if (FMT == FMT_F8) {
                    const int C = 2 * tt + (u >> 1);  // 16-k chunk index
                    du = 2 * (C >> 2) + (u & 1), dt = C & 3;
                }
                dstx[((s * UNITS + du) * MX + m) * 4 + dt] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// NT: 8-token column tiles ; MXS: token rows staged in shared memory when NT == 1 (1, 2, 4 or 8: the MMA's 8 columns replicate them, so
// a 2-token step stages a quarter of the activations of an 8-token one) ; RT: 16-row tiles per warp
// XG: the activations come pre-staged from global memory (kf_gemv_xprep_kernel) instead of being staged per CTA in shared memory.  With 4..8
// tokens the staging area (2 KB per k-step) otherwise caps the k-slice at 18-28 steps, i.e. forces 2-3 k-splits and as many waves on the
// large shapes, and every CTA repeats the fp16 conversion of its slice; pre-staged, the slice length is free (one wave) and the
// conversion happens once.  The fragment loads hit L1 (every warp of every CTA on the SM reads the same 2 KB per k-step).
template <int FMT, int MODE, int NT, int MXS, int RT, int XG = 0>
#ifndef KF_GEMV_OCC
#define KF_GEMV_OCC 3
#endif
__global__ void __launch_bounds__(kThreads, (NT >= 4 ? 1 : NT == 1 ? KF_GEMV_OCC : 2)) kf_gemv_kernel(const GemvParams p) {
    using F = Fmt<FMT>;
    constexpr int DEPTH = RT == 2 ? F::D2 : F::D1, CPB = F::CPB, NCH = F::NCH;
    constexpr int TB = fmt_tb(F::BITS), NR = TB / 4;  // bytes / 32-bit registers per thread, row and k-step
    constexpr int STEPB = KSTEP * F::BITS / 8;         // bytes per row per k-step
    constexpr int MX = NT == 1 ? MXS : 8 * NT;         // token rows staged in shared memory
    constexpr bool M1 = MX == 1;
    constexpr int MP = 8 * NT;                         // token columns of the fp32 tile
    constexpr int ROWS = 128 * RT, HALF = ROWS / 2, WROWS = 16 * RT, TS = ROWS + 4, GS = ROWS + 1;
    static_assert(TB == CPB * NCH, "ring chunking");
    extern __shared__ uint4 smem[];
    __shared__ int s_last;
    __shared__ float s_red[32];
    __shared__ float s_scale[64];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int rb = blockIdx.x, split = blockIdx.y;
    const bool swiglu = p.epilogue == EPI_SWIGLU;
    kf_grid_launch_dependents();  // the next kernel of the stream may start its own weight prefetch as soon as all our CTAs are running
    if (NT == 1 && p.cluster) cluster_arrive();  // matched by the wait in front of the first remote store: by then every CTA of the cluster runs

    // ---- which rows does this CTA / warp own? -------------------------------------------------------------------------------
    int segi = 0;
    if (!swiglu) {
        if (p.nseg > 1 && rb >= p.seg[1].rb0) segi = 1;
        if (p.nseg > 2 && rb >= p.seg[2].rb0) segi = 2;
    }
    const GemvSeg& sgw = p.seg[swiglu ? (warp >> 2) : segi];  // segment of this warp
    const int wrow0    = swiglu ? rb * HALF + (warp & 3) * WROWS : (rb - sgw.rb0) * ROWS + warp * WROWS;

    const int s_begin = (int)(((long long)split * p.steps_total) / p.S);
    const int s_end   = (int)(((long long)(split + 1) * p.steps_total) / p.S);
    const int nsteps  = s_end - s_begin;
    uint32_t* sgam    = reinterpret_cast<uint32_t*>(smem + (XG ? (size_t)0 : (size_t)p.nsteps_max * UNITS * MX * 4));
    // per-thread ring of packed weight words: [DEPTH][RT][2 (row g / g+8)][NCH][256 threads] x CPB bytes
    uint8_t* ring = reinterpret_cast<uint8_t*>(smem) + p.ring_off;

    // ---- weight stream: start the first DEPTH k-steps NOW -- they do not depend on the activations, so the copies overlap the
    //      whole prologue (norm, activation staging, gama staging) --------------------------------------------------------------
    const bool wactive     = nsteps > 0 && wrow0 < sgw.rows;
    const size_t row_bytes = (size_t)p.K * F::BITS / 8;
    const uint8_t* pa[RT];
    {
        int toff;  // byte offset of this thread's slot inside one k-step of a row
        if (FMT == FMT_Q2)
            toff = 16 * (t >> 1) + 8 * (1 - (t & 1));  // word.high holds the first 32 codes (PackedQ.hpp:185-198)
        else if (FMT == FMT_Q1)
            toff = 12 - 4 * t;                          // high.hi32 holds codes 0..31 (PackedQ.hpp:200-211)
        else
            toff = 16 * t;  // 16-byte formats: chunk ch of thread t sits at (4*ch + t)*16, so a quad always covers 64 contiguous bytes
#pragma unroll
        for (int rt = 0; rt < RT; rt++) {
            int r0 = wrow0 + rt * 16;
            if (r0 + 16 > sgw.rows) r0 = sgw.rows - 16;  // out-of-range tile: read valid rows, results are dropped by the epilogue
            pa[rt] = sgw.data + (size_t)(r0 + g) * row_bytes + (size_t)s_begin * STEPB + toff;
        }
    }
    auto ring_at = [&](int slot_, int rt, int half, int ch) -> uint8_t* {
        return ring + ((size_t)(((slot_ * RT + rt) * 2 + half) * NCH + ch) * kThreads + tid) * CPB;
    };
    auto issue_stage = [&](int slot_, int sl) {
#pragma unroll
        for (int rt = 0; rt < RT; rt++)
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    cp_async<CPB>(ring_at(slot_, rt, half, ch), pa[rt] + (size_t)half * 8 * row_bytes + (size_t)sl * STEPB + ch * 64);
    };
    if (wactive) {
#pragma unroll
        for (int i = 0; i < DEPTH; i++) {
            if (i < nsteps) issue_stage(i, i);
            cp_async_commit();
        }
    }

    // ---- stage zero/step of every (row, k-step) of the CTA tile.  One thread per (row, array): a contiguous run of nsteps bf16,
    //      fetched with 16-byte loads that are all issued before the first use (one DRAM round trip, not one per element) -----------
    if (MODE != MODE_PLAIN && nsteps > 0) {
        const int gpr = (p.K >> 7) >> p.gshift;  // groups per row
        for (int idx = tid; idx < 2 * ROWS; idx += kThreads) {
            const int r = idx % ROWS, which = idx / ROWS;  // which: 0 = zero, 1 = step
            const GemvSeg& sr = p.seg[swiglu ? (r >= HALF) : segi];
            const int grow    = swiglu ? rb * HALF + (r & (HALF - 1)) : (rb - sr.rb0) * ROWS + r;
            uint16_t* dst     = reinterpret_cast<uint16_t*>(sgam + r) + which;  // low half = zero, high half = step
            if (grow >= sr.rows) {
                for (int s = 0; s < nsteps; s++) dst[(size_t)s * GS * 2] = 0;
                continue;
            }
            const uint16_t* src = (which ? sr.step : sr.zero) + (size_t)grow * gpr;
            if (p.gshift == 0 && ((reinterpret_cast<uintptr_t>(src + s_begin) & 15) == 0)) {
                const uint4* v4 = reinterpret_cast<const uint4*>(src + s_begin);
                int s = 0;
                for (; s + 32 <= nsteps; s += 32) {  // 4 x 16 bytes in flight
                    uint4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = __ldg(v4 + (s >> 3) + j);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t q[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                        for (int e = 0; e < 8; e++) dst[(size_t)(s + 8 * j + e) * GS * 2] = (uint16_t)(q[e >> 1] >> ((e & 1) * 16));
                    }
                }
                for (; s + 8 <= nsteps; s += 8) {
                    const uint4 v = __ldg(v4 + (s >> 3));
                    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 8; e++) dst[(size_t)(s + e) * GS * 2] = (uint16_t)(q[e >> 1] >> ((e & 1) * 16));
                }
                for (; s < nsteps; s++) dst[(size_t)s * GS * 2] = __ldg(src + s_begin + s);
            } else {
                int s = 0;
                for (; s + 8 <= nsteps; s += 8) {
                    uint16_t v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = __ldg(src + ((s_begin + s + e) >> p.gshift));
#pragma unroll
                    for (int e = 0; e < 8; e++) dst[(size_t)(s + e) * GS * 2] = v[e];
                }
                for (; s < nsteps; s++) dst[(size_t)s * GS * 2] = __ldg(src + ((s_begin + s) >> p.gshift));
            }
        }
    }

    // ---- everything above touched only the weights.  From here on the kernel reads what its predecessor in the stream wrote (x,
    //      residual) and writes buffers the predecessor may still be using (split-K workspace, y): wait for it to finish.  Under
    //      programmatic dependent launch this CTA may have been running for a while already, with its weight stream in flight -------
    kf_grid_dependency_wait();

    // ---- optional fused RMSNorm + the activations of this k-slice into shared memory, permuted to the fragment order ------------------
    if constexpr (!XG)
        gemv_stage_x<FMT, MODE, MX>(smem, reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(smem) + p.sx_off), p, s_begin, nsteps, s_red, s_scale);
    __syncthreads();
    float* sxs = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem) + p.sx_off);  // [nsteps][MX] group sums of the activations
    if (MODE == MODE_FACTOR) {
        for (int it = tid; it < nsteps * MX; it += kThreads) {
            const int s = it / MX, m = it - s * MX;
            float sum = 0.f;
            for (int u = 0; u < UNITS; u++)
#pragma unroll
                for (int tt = 0; tt < 4; tt++) {
                    const uint4 v = smem[((s * UNITS + u) * MX + m) * 4 + tt];
                    sum += bf16lo(v.x), sum += bf16hi(v.x), sum += bf16lo(v.y), sum += bf16hi(v.y);
                    sum += bf16lo(v.z), sum += bf16hi(v.z), sum += bf16lo(v.w), sum += bf16hi(v.w);
                }
            sxs[it] = sum;
        }
        __syncthreads();
    }

    float acc[RT][NT][4];
#pragma unroll
    for (int rt = 0; rt < RT; rt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[rt][nt][j] = 0.f;

    if (wactive) {
        const uint32_t bias2 = pack_bf16x2((float)(128 + p.qbias), (float)(128 + p.qbias));
        const float fqb      = (float)p.qbias;
        const int xlane       = (MX >= 8 ? g : (g & (MX - 1))) * 4 + t;  // column g of the MMA reads token g mod MX
        const uint32_t* gbase = sgam + warp * WROWS + g;
        int slot = 0;
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            cp_async_wait<DEPTH - 1>();  // this thread's copies of k-step s have landed (each thread reads only what it copied)
            uint32_t wreg[RT][2][NR];
#pragma unroll
            for (int rt = 0; rt < RT; rt++)
#pragma unroll
                for (int half = 0; half < 2; half++)
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) {
                        if constexpr (CPB == 16) {
                            const uint4 v = *reinterpret_cast<const uint4*>(ring_at(slot, rt, half, ch));
                            wreg[rt][half][4 * ch + 0] = v.x, wreg[rt][half][4 * ch + 1] = v.y;
                            wreg[rt][half][4 * ch + 2] = v.z, wreg[rt][half][4 * ch + 3] = v.w;
                        } else if constexpr (CPB == 8) {
                            const uint2 v = *reinterpret_cast<const uint2*>(ring_at(slot, rt, half, ch));
                            wreg[rt][half][0] = v.x, wreg[rt][half][1] = v.y;
                        } else {
                            wreg[rt][half][0] = *reinterpret_cast<const uint32_t*>(ring_at(slot, rt, half, ch));
                        }
                    }
            uint32_t gm[RT][6];
            float fstep[RT][2];
            if (MODE != MODE_PLAIN && MODE != MODE_FACTOR && MODE != MODE_FAST) {
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    const uint32_t ga = gbase[s * GS + rt * 16], gb = gbase[s * GS + rt * 16 + 8];
                    gm[rt][0] = __byte_perm(ga, 0u, 0x3232), gm[rt][1] = __byte_perm(ga, 0u, 0x1010);
                    gm[rt][3] = __byte_perm(gb, 0u, 0x3232), gm[rt][4] = __byte_perm(gb, 0u, 0x1010);
                    gm[rt][2] = gm[rt][5] = 0u;
                    if (MODE == MODE_AFFINE_FMA) gm[rt][1] ^= 0x80008000u, gm[rt][4] ^= 0x80008000u;  // -zero
                    if (MODE == MODE_AFFINE) {
                        gm[rt][2] = bf162_as_u32(__hmul2_rn(u32_as_bf162(gm[rt][0]), u32_as_bf162(0xC300C300u)));  // -128*step, exact
                        gm[rt][5] = bf162_as_u32(__hmul2_rn(u32_as_bf162(gm[rt][3]), u32_as_bf162(0xC300C300u)));
                    }
                    fstep[rt][0] = bf16hi(ga), fstep[rt][1] = bf16hi(gb);
                }
            }
            float accg[RT][NT][4];
            if (MODE == MODE_SCALE || MODE == MODE_FACTOR) {
#pragma unroll
                for (int rt = 0; rt < RT; rt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int j = 0; j < 4; j++) accg[rt][nt][j] = 0.f;
            }
            float4 sv[NT][2];  // MODE_FAST: {-, 2^-24 Sx, 2^24 / scale} of this k-step's group for the thread's two token columns
            if (MODE == MODE_FAST) {
                const float4* sx4 = XG ? p.sxg + (size_t)(s_begin + s) * MX
                                       : reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(smem) + p.sx_off) + s * MX;
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    const int i0 = MX >= 8 ? nt * 8 + 2 * t : ((2 * t) & (MX - 1)), i1 = MX >= 8 ? nt * 8 + 2 * t + 1 : ((2 * t + 1) & (MX - 1));
                    if constexpr (XG) {
                        sv[nt][0] = __ldg(sx4 + i0), sv[nt][1] = __ldg(sx4 + i1);
                    } else {
                        sv[nt][0] = sx4[i0], sv[nt][1] = sx4[i1];
                    }
#pragma unroll
                    for (int rt = 0; rt < RT; rt++)
#pragma unroll
                        for (int j = 0; j < 4; j++) accg[rt][nt][j] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UNITS; u++) {
                uint4 xb[NT];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    if constexpr (XG)
                        xb[nt] = __ldg(p.xg + (((size_t)(s_begin + s) * UNITS + u) * MX + (MX >= 8 ? nt * 8 : 0)) * 4 + xlane);
                    else
                        xb[nt] = smem[((s * UNITS + u) * MX + (MX >= 8 ? nt * 8 : 0)) * 4 + xlane];
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
#pragma unroll
                    for (int rt = 0; rt < RT; rt++) {
                        uint32_t a[4];
                        build_a<FMT, MODE, NR>(a, wreg[rt][0], wreg[rt][1], u, h, gm[rt], bias2, p.lop_mask, p.lop_magic);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            const uint32_t b0 = h ? xb[nt].z : xb[nt].x, b1 = h ? xb[nt].w : xb[nt].y;
                            if (MODE == MODE_FAST)
                                mma_f16_16816(accg[rt][nt], a, b0, b1);
                            else if (MODE == MODE_SCALE || MODE == MODE_FACTOR)
                                mma_bf16_16816(accg[rt][nt], a, b0, b1);
                            else
                                mma_bf16_16816(acc[rt][nt], a, b0, b1);
                        }
                    }
                }
            }
            if (MODE == MODE_SCALE) {
#pragma unroll
                for (int rt = 0; rt < RT; rt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        acc[rt][nt][0] = fmaf(fstep[rt][0], accg[rt][nt][0], acc[rt][nt][0]);
                        acc[rt][nt][1] = fmaf(fstep[rt][0], accg[rt][nt][1], acc[rt][nt][1]);
                        acc[rt][nt][2] = fmaf(fstep[rt][1], accg[rt][nt][2], acc[rt][nt][2]);
                        acc[rt][nt][3] = fmaf(fstep[rt][1], accg[rt][nt][3], acc[rt][nt][3]);
                    }
            }
            if (MODE == MODE_FAST) {  // y += 2^(24 - shift) * (step * sum(c x') - (zero + qbias step) * Sx')
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    const uint32_t ga = gbase[s * GS + rt * 16], gb = gbase[s * GS + rt * 16 + 8];
                    const float sa = bf16hi(ga), sb = bf16hi(gb);
                    const float za = fmaf(fqb, sa, bf16lo(ga)), zb = fmaf(fqb, sb, bf16lo(gb));
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        acc[rt][nt][0] = fmaf(sv[nt][0].z, fmaf(sa, accg[rt][nt][0], -za * sv[nt][0].y), acc[rt][nt][0]);
                        acc[rt][nt][1] = fmaf(sv[nt][1].z, fmaf(sa, accg[rt][nt][1], -za * sv[nt][1].y), acc[rt][nt][1]);
                        acc[rt][nt][2] = fmaf(sv[nt][0].z, fmaf(sb, accg[rt][nt][2], -zb * sv[nt][0].y), acc[rt][nt][2]);
                        acc[rt][nt][3] = fmaf(sv[nt][1].z, fmaf(sb, accg[rt][nt][3], -zb * sv[nt][1].y), acc[rt][nt][3]);
                    }
                }
            }
            if (MODE == MODE_FACTOR) {
                const float off = (float)(128 + p.qbias);
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    const uint32_t ga = gbase[s * GS + rt * 16], gb = gbase[s * GS + rt * 16 + 8];
                    const float sa = bf16hi(ga), sb = bf16hi(gb);
                    const float ka = fmaf(off, sa, bf16lo(ga)), kb = fmaf(off, sb, bf16lo(gb));  // (128+qbias)*step + zero
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        const float2 sx = MX >= 8 ? *reinterpret_cast<const float2*>(sxs + s * MX + nt * 8 + 2 * t)
                                                  : make_float2(sxs[s * MX + ((2 * t) & (MX - 1))], sxs[s * MX + ((2 * t + 1) & (MX - 1))]);
                        acc[rt][nt][0] += fmaf(sa, accg[rt][nt][0], -ka * sx.x);
                        acc[rt][nt][1] += fmaf(sa, accg[rt][nt][1], -ka * sx.y);
                        acc[rt][nt][2] += fmaf(sb, accg[rt][nt][2], -kb * sx.x);
                        acc[rt][nt][3] += fmaf(sb, accg[rt][nt][3], -kb * sx.y);
                    }
                }
            }
            // the slot has been consumed (its registers fed the MMAs above): refill it DEPTH steps ahead
            if (s + DEPTH < nsteps) issue_stage(slot, s + DEPTH);
            cp_async_commit();  // one group per k-step, empty at the tail, so the wait count stays uniform
            slot = slot + 1 == DEPTH ? 0 : slot + 1;
        }
    }
    cp_async_wait<0>();

    // ---- scatter fragments to the fp32 tile [MP][ROWS] in shared memory ---------------------------------------------------------
    __syncthreads();  // everyone is done reading x / gama / ring
    float* tile = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int rt = 0; rt < RT; rt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const int m0 = nt * 8 + 2 * t, r = warp * WROWS + rt * 16 + g;
            tile[(m0 + 0) * TS + r]     = acc[rt][nt][0];
            tile[(m0 + 1) * TS + r]     = acc[rt][nt][1];
            tile[(m0 + 0) * TS + r + 8] = acc[rt][nt][2];
            tile[(m0 + 1) * TS + r + 8] = acc[rt][nt][3];
        }
    __syncthreads();

    // ---- split-K: publish the partial tile; the last CTA of this row block reduces in fixed order ---------------------------
    if (NT == 1 && p.cluster) {
        // the k-slices of this row block are the CTAs of one cluster (up to 8 tokens): every slice stores its M x ROWS partial sums into the
        // leader's shared memory, the leader adds them in slice order (deterministic).  No global workspace, no atomics, no second
        // DRAM/L2 round trip.
        float* red  = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem) + p.red_off);
        const int n = p.M * ROWS;
        cluster_wait();
        for (int e = tid; e < n; e += kThreads) st_cluster_f32(red + (size_t)split * n + e, 0, tile[(e / ROWS) * TS + (e % ROWS)]);
        cluster_arrive();
        cluster_wait();
        if (split != 0) return;
        for (int e = tid; e < n; e += kThreads) {
            float sum = 0.f;
            for (int sp = 0; sp < p.S; sp++) sum += red[(size_t)sp * n + e];
            tile[(e / ROWS) * TS + (e % ROWS)] = sum;
        }
        __syncthreads();
    } else if (p.S > 1) {
        float* wsp = p.ws + ((size_t)split * p.total_rb + rb) * (size_t)(MP * ROWS);
        for (int e = tid; e < p.M * ROWS; e += kThreads) {
            const int m = e / ROWS, r = e % ROWS;
            __stcg(wsp + m * ROWS + r, tile[m * TS + r]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(p.cnt + rb, 1u);
            s_last              = (prev == (unsigned)(p.S - 1));
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int e = tid; e < p.M * ROWS; e += kThreads) {
            const int m = e / ROWS, r = e % ROWS;
            float sum = 0.f;
            for (int sp = 0; sp < p.S; sp++) sum += __ldcg(p.ws + ((size_t)sp * p.total_rb + rb) * (size_t)(MP * ROWS) + m * ROWS + r);
            tile[m * TS + r] = sum;
        }
        if (tid == 0) p.cnt[rb] = 0u;  // self-reset for the next launch
        __syncthreads();
    }

    // ---- epilogue -----------------------------------------------------------------------------------------------------------
    if (p.epilogue == EPI_TP) {
        // fused tensor-parallel exchange (kf_tp.cuh): push the fp32 rows of this tile into slot [rank] of EVERY peer, then add the `world`
        // partials of the row block in rank order (every rank does, with identical results), apply the two roundings of the single-GPU
        // epilogue (bf16 of the matmul, bf16 of residual + that) and write the rows of y -- the next kernel of the stream reads plain
        // activations and knows nothing of the exchange.  Flag-in-data: each 8 bytes that cross NVLink carry their epoch.
        const GemvSeg& sg = p.seg[0];
        const int E = sg.rows, rbase = rb * ROWS, W = p.tp.world, me = p.tp.rank;
        const unsigned e = kftp::epoch(p.tp, p.tp_out), par = e & 1u;
        constexpr int HP = ROWS / 2;  // row pairs of the tile
        for (int idx = tid; idx < p.M * HP; idx += kThreads) {
            const int m = idx / HP, r = (idx % HP) * 2, row = rbase + r;
            if (row >= E) continue;
            const size_t off   = ((size_t)m * E + row) * 8;
            const uint32_t v0 = __float_as_uint(tile[m * TS + r]), v1 = __float_as_uint(tile[m * TS + r + 1]);
            for (int w = 1; w < W; w++) {
                const int peer = me + w < W ? me + w : me + w - W;  // everybody starts at a different peer
                kftp::st16_sys(kftp::scat(p.tp, peer, par, me) + off, v0, e, v1, e);
            }
        }
        kftp::SpinGuard sgd;
        for (int idx = tid; idx < p.M * HP; idx += kThreads) {
            const int m = idx / HP, r = (idx % HP) * 2, row = rbase + r;
            if (row >= E) continue;
            const size_t el = (size_t)m * E + row;
            uint4 v[KF_TP_MAX_WORLD];
#pragma unroll
            for (int w = 0; w < KF_TP_MAX_WORLD; w++)
                if (w < W && w != me) v[w] = kftp::ld16_sys(kftp::scat(p.tp, me, par, w) + el * 8);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w = 0; w < KF_TP_MAX_WORLD; w++) {
                if (w >= W) break;
                if (w == me) {
                    a0 += tile[m * TS + r], a1 += tile[m * TS + r + 1];
                    continue;
                }
                while (v[w].y != e || v[w].w != e) {
                    sgd.tick();
                    v[w] = kftp::ld16_sys(kftp::scat(p.tp, me, par, w) + el * 8);
                }
                a0 += __uint_as_float(v[w].x), a1 += __uint_as_float(v[w].z);
            }
            const uint32_t rp = *reinterpret_cast<const uint32_t*>(p.residual + el);
            *reinterpret_cast<uint32_t*>(sg.y + el) =
                pack_bf16x2(bf16lo(rp) + bf16_bits_to_f32(f32_to_bf16_bits(a0)), bf16hi(rp) + bf16_bits_to_f32(f32_to_bf16_bits(a1)));
        }
        return;
    }
    if (!swiglu) {
        const GemvSeg& sg = p.seg[segi];
        const int rbase   = (rb - sg.rb0) * ROWS;
        for (int e = tid; e < p.M * ROWS; e += kThreads) {
            const int m = e / ROWS, r = e % ROWS, row = rbase + r;
            if (row >= sg.rows) continue;
            if (p.epilogue == EPI_F32) {  // tensor-parallel partial sums stay fp32 until the all-reduce
                reinterpret_cast<float*>(sg.y)[(size_t)m * sg.rows + row] = tile[m * TS + r];
                continue;
            }
            uint16_t v = f32_to_bf16_bits(tile[m * TS + r]);  // the reference's GEMM writes bf16 (gemm.cu:124-126)
            if (p.epilogue == EPI_RESIDUAL)                   // then CU_add3 adds the residual in fp32 (packedN.cuh:867-875)
                v = f32_to_bf16_bits(bf16_bits_to_f32(p.residual[(size_t)m * sg.rows + row]) + bf16_bits_to_f32(v));
            sg.y[(size_t)m * sg.rows + row] = v;
        }
    } else {
        const int rows = p.seg[0].rows;
        for (int e = tid; e < p.M * HALF; e += kThreads) {
            const int m = e / HALF, r = e % HALF, row = rb * HALF + r;
            if (row >= rows) continue;
            const float gt = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * TS + r]));
            const float up = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * TS + HALF + r]));
            p.seg[0].y[(size_t)m * rows + row] = f32_to_bf16_bits((gt * up) / (1.0f + expf(-gt)));  // CU_swiglu_v0, Activation.cu:86-93
        }
    }
}

// RMSNorm (optional) + fragment-order fp16 staging of ALL of x, once, for the XG variants of the kernel above: grid = ceil(k-steps / prep_steps)
template <int FMT, int MX>
__global__ void __launch_bounds__(kThreads) kf_gemv_xprep_kernel(const GemvParams p) {
    __shared__ float s_red[32];
    __shared__ float s_scale[64];
    kf_grid_launch_dependents();  // the matmul behind us may start streaming its weights
    kf_grid_dependency_wait();    // x comes from the kernel before us
    const int s0 = blockIdx.x * p.prep_steps, n = min(p.prep_steps, p.steps_total - s0);
    gemv_stage_x<FMT, MODE_FAST, MX>(p.xg_out + (size_t)s0 * UNITS * MX * 4, p.sxg_out + (size_t)s0 * MX, p, s0, n, s_red, s_scale);
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
constexpr size_t kSmemCap  = 100 * 1024;  // two CTAs per SM
#ifndef KF_GEMV_SOFT_KB
#define KF_GEMV_SOFT_KB (KF_GEMV_OCC == 4 ? 55 : 73)
#endif
constexpr size_t kSmemSoft = (size_t)KF_GEMV_SOFT_KB * 1024;  // KF_GEMV_OCC CTAs per SM

struct FmtInfo {
    int bits, cpb, nch, d1, d2;
};
static FmtInfo fmt_info(int fmt) {
    switch (fmt) {
        case FMT_Q4: return {Fmt<FMT_Q4>::BITS, Fmt<FMT_Q4>::CPB, Fmt<FMT_Q4>::NCH, Fmt<FMT_Q4>::D1, Fmt<FMT_Q4>::D2};
        case FMT_Q2: return {Fmt<FMT_Q2>::BITS, Fmt<FMT_Q2>::CPB, Fmt<FMT_Q2>::NCH, Fmt<FMT_Q2>::D1, Fmt<FMT_Q2>::D2};
        case FMT_Q1: return {Fmt<FMT_Q1>::BITS, Fmt<FMT_Q1>::CPB, Fmt<FMT_Q1>::NCH, Fmt<FMT_Q1>::D1, Fmt<FMT_Q1>::D2};
        case FMT_F8: return {Fmt<FMT_F8>::BITS, Fmt<FMT_F8>::CPB, Fmt<FMT_F8>::NCH, Fmt<FMT_F8>::D1, Fmt<FMT_F8>::D2};
        default: return {Fmt<FMT_BF16>::BITS, Fmt<FMT_BF16>::CPB, Fmt<FMT_BF16>::NCH, Fmt<FMT_BF16>::D1, Fmt<FMT_BF16>::D2};
    }
}
static size_t ring_bytes(int fmt, int rt) {
    const FmtInfo f = fmt_info(fmt);
    return (size_t)(rt == 2 ? f.d2 : f.d1) * rt * 2 * f.nch * kThreads * f.cpb;
}
// shared bytes per k-step: activations + gama
static size_t step_bytes(int mode, int mx, int rows_cta, bool xg = false) {
    if (xg) return (size_t)(rows_cta + 1) * 4;  // activations and group sums live in global memory: only zero / step are staged
    return (size_t)UNITS * mx * 64 + (mode == MODE_PLAIN ? 0 : (size_t)(rows_cta + 1) * 4) + (mode == MODE_FACTOR ? (size_t)mx * 4 : 0) +
           (mode == MODE_FAST ? (size_t)mx * 16 : 0);
}

template <int FMT, int MODE, int NT, int MXS, int RT, int XG = 0>
int launch_one(kf_ctx* ctx, const GemvParams& p0) {
    constexpr int MX = NT == 1 ? MXS : 8 * NT, MP = 8 * NT, ROWS = 128 * RT;
    GemvParams p     = p0;
    size_t head      = (size_t)p.nsteps_max * step_bytes(MODE, MX, ROWS, XG != 0);
    size_t tilebytes = (size_t)MP * (ROWS + 4) * 4;
    p.ring_off       = (int)((head + 15) & ~(size_t)15);
    p.sx_off         = (int)(p.ring_off + ring_bytes(FMT, RT));
    size_t sxbytes   = XG ? 0 : MODE == MODE_FACTOR ? (size_t)p.nsteps_max * MX * 4 : MODE == MODE_FAST ? (size_t)p.nsteps_max * MX * 16 : 0;
    size_t smem      = std::max((size_t)p.sx_off + sxbytes, tilebytes);
    if (p.cluster) {
        p.red_off = (int)((smem + 15) & ~(size_t)15);
        smem      = (size_t)p.red_off + (size_t)p.S * p.M * ROWS * 4;
        if (NT != 1 || smem > kSmemCap) {  // no room for the merge area: global workspace + last-CTA reduction
            p.cluster = 0, smem = std::max((size_t)p.sx_off + sxbytes, tilebytes);
            int rc = kf_ensure_gemv_ws(ctx, (size_t)p.S * p.total_rb * MP * ROWS * sizeof(float), p.total_rb);
            if (rc) return rc;
            p.ws = ctx->gemv_ws, p.cnt = ctx->gemv_cnt;
        }
    }
    KF_REQUIRE(ctx, smem <= kSmemCap, "internal: k-slice does not fit shared memory");
    auto kern            = kf_gemv_kernel<FMT, MODE, NT, MXS, RT, XG>;
    static bool attr_set[kf_ctx::kMaxDevices] = {};  // function attributes are per device (one flag per instantiation and device)
    if (!attr_set[ctx->device]) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
        attr_set[ctx->device] = true;
    }
    dim3 grid(p.total_rb, p.S);
    if (p.cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid, cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id               = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = (unsigned)p.S, attr[0].val.clusterDim.z = 1;
        attr[1].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = ctx->pdl ? 2 : 1;
        KF_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    } else {
        KF_CUDA(ctx, kf_launch_pdl(ctx, kern, grid, dim3(kThreads), smem, p));
    }
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

template <int FMT, int MODE>
int launch_nt(kf_ctx* ctx, const GemvParams& p, int rt) {
    if constexpr (MODE == MODE_FAST) {
        if (p.xg && rt == 1) {  // pre-staged activations (4..8 tokens): stage them once, then the matmul
            constexpr int PS = 8;
            GemvParams q = p;
            q.prep_steps = PS;
            const dim3 grid((p.steps_total + PS - 1) / PS);
            if (p.M <= 4)
                KF_CUDA(ctx, kf_launch_pdl(ctx, kf_gemv_xprep_kernel<FMT, 4>, grid, dim3(kThreads), 0, q));
            else
                KF_CUDA(ctx, kf_launch_pdl(ctx, kf_gemv_xprep_kernel<FMT, 8>, grid, dim3(kThreads), 0, q));
            KF_LAUNCH_CHECK(ctx);
            return p.M <= 4 ? launch_one<FMT, MODE, 1, 4, 1, 1>(ctx, p) : launch_one<FMT, MODE, 1, 8, 1, 1>(ctx, p);
        }
    }
    if (p.M == 1) return rt == 2 ? launch_one<FMT, MODE, 1, 1, 2>(ctx, p) : launch_one<FMT, MODE, 1, 1, 1>(ctx, p);
    if (p.M == 2 && rt != 2) return launch_one<FMT, MODE, 1, 2, 1>(ctx, p);
    if (p.M <= 4 && rt != 2) return launch_one<FMT, MODE, 1, 4, 1>(ctx, p);
    if (p.M <= 8) return rt == 2 ? launch_one<FMT, MODE, 1, 8, 2>(ctx, p) : launch_one<FMT, MODE, 1, 8, 1>(ctx, p);
    if (p.M <= 16) return launch_one<FMT, MODE, 2, 8, 1>(ctx, p);
    if (p.M <= 32) return launch_one<FMT, MODE, 4, 8, 1>(ctx, p);
    return launch_one<FMT, MODE, 8, 8, 1>(ctx, p);
}

int gemv_dispatch(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                  const void* norm_w, float norm_eps, int xid_out = -1) {
    KF_REQUIRE(ctx, n >= 1 && n <= 3 && M >= 1 && M <= 64 && x, "1..3 weights, 1..64 tokens");
    const int type = w[0].type, K = w[0].cols;
    int fmt, mode;
    switch (type) {
        case KF_T_BF16: fmt = FMT_BF16, mode = MODE_PLAIN; break;
        case KF_T_F8E5M2: fmt = FMT_F8, mode = MODE_PLAIN; break;
        case KF_T_Q4: fmt = FMT_Q4, mode = !ctx->gemv_exact ? MODE_FAST : ctx->deq_fma ? MODE_AFFINE_FMA : w[0].qbias == 0 ? MODE_AFFINE : MODE_AFFINE_SYM; break;
        case KF_T_Q2: fmt = FMT_Q2, mode = !ctx->gemv_exact ? MODE_FAST : ctx->deq_fma ? MODE_AFFINE_FMA : w[0].qbias == 0 ? MODE_AFFINE : MODE_AFFINE_SYM; break;
        case KF_T_SIGN: fmt = FMT_Q2, mode = !ctx->gemv_exact ? MODE_FAST : MODE_SCALE; break;
        case KF_T_BINARY: fmt = FMT_Q1, mode = !ctx->gemv_exact ? MODE_FAST : MODE_SCALE; break;
        default: return KF_ERR_UNSUPPORTED;
    }
    KF_REQUIRE(ctx, K % KSTEP == 0, "K must be a multiple of 128");
    GemvParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = n, p.x = (const uint16_t*)x, p.residual = (const uint16_t*)residual, p.M = M, p.K = K;
    p.norm_w = (const uint16_t*)norm_w, p.norm_eps = norm_eps;
    p.steps_total = K / KSTEP, p.qbias = w[0].qbias, p.epilogue = epilogue;
    p.lop_mask = fmt == FMT_Q4 ? 0x000F000Fu : fmt == FMT_Q2 ? 0x00030003u : 0x00010001u;
    p.lop_magic = 0x43004300u;  // bf16x2 128.0
    p.tp_out = -1;
    if (xid_out >= 0) {
        KF_REQUIRE(ctx, kf_tp_view(ctx, &p.tp) == KF_OK, "fused exchange: peer buffers not attached");
        KF_REQUIRE(ctx, M <= 8 && xid_out < p.tp.stride && n == 1 && epilogue == EPI_TP && residual && (size_t)M * w[0].rows <= KF_TP_LL_ELEMS,
                   "fused exchange: one weight, <= 8 tokens, M x rows within a slot, a residual");
        p.tp_out = xid_out;
    }
    KF_REQUIRE(ctx, (epilogue == EPI_TP) == (p.tp_out >= 0), "fused exchange epilogue");
    int total_rows = 0;
    for (int i = 0; i < n; i++) {
        KF_REQUIRE(ctx, w[i].type == type && w[i].cols == K && w[i].qbias == w[0].qbias && w[i].group == w[0].group,
                   "fused weights must share type / K / quant card");
        KF_REQUIRE(ctx, w[i].rows % 16 == 0 && w[i].rows >= 16 && w[i].data_dev && y[i], "rows must be a multiple of 16");
        if (mode != MODE_PLAIN)
            KF_REQUIRE(ctx, kf_has_gama(w[i]) && w[i].group >= 128 && (w[i].group & (w[i].group - 1)) == 0 && K % w[i].group == 0,
                       "fused path needs group = 128 * 2^n dividing K");
        total_rows += w[i].rows;
    }
    if (mode != MODE_PLAIN) {
        int gs = 0;
        while ((128 << gs) < w[0].group) gs++;
        p.gshift = gs;
    }
    if (epilogue == EPI_SWIGLU) KF_REQUIRE(ctx, n == 2 && w[0].rows == w[1].rows, "swiglu needs gate and up of equal shape");
    if (epilogue == EPI_RESIDUAL) KF_REQUIRE(ctx, n == 1 && residual, "residual epilogue takes one weight");

    // ---- tile shape: 32 rows per warp (RT = 2) when there are enough rows to keep every SM busy, else 16 -------------------------
    int rt = 1;
    if (M <= 8 && fmt != FMT_BF16 && fmt != FMT_F8) {
        rt = 1;  // measured (profiles/r01_gemv_sweep_v4.txt): 16 rows per warp is never slower than 32 on B200
        (void)total_rows;
        if (ctx->gemv_variant == 2) rt = 2;
    }
    const int rows_cta = 128 * rt;
    int rb = 0;
    for (int i = 0; i < n; i++) {
        p.seg[i].data = (const uint8_t*)w[i].data_dev, p.seg[i].y = (uint16_t*)y[i], p.seg[i].rows = w[i].rows, p.seg[i].rb0 = rb;
        if (mode != MODE_PLAIN) p.seg[i].zero = kf_gama_zero(w[i]), p.seg[i].step = kf_gama_step(w[i]);
        rb += (w[i].rows + rows_cta - 1) / rows_cta;
    }
    if (epilogue == EPI_SWIGLU) rb = (w[0].rows + rows_cta / 2 - 1) / (rows_cta / 2);
    p.total_rb = rb;

    // ---- k-split: pick the S that minimises  waves(S) x (fixed CTA cost + k-steps per CTA)  under the shared-memory budget -------
    const int MXs      = M == 1 ? 1 : M == 2 ? 2 : M <= 4 ? 4 : (M <= 8 ? 8 : M <= 16 ? 16 : M <= 32 ? 32 : 64);
    // Plan the k-split for a given staging of the activations; returns the number of waves of the chosen plan.
    // Measured on B200 (profiles/r02_gemv_splitk.jsonl): a k-slice advances at ~0.55 us per k-step whatever the occupancy until the launch
    // as a whole saturates HBM.  So: ONE wave, as many slices as fit into it; and when the slices of a row block merge inside a thread-block
    // cluster (single token, S <= 8) only the cluster sizes that tile the GPCs (1, 2, 4, 8): S = 5 on 10240 x 5120 is 13.6 us where S = 4 is
    // 10.3.  The merge through the global workspace (S > 8) costs ~3.5 us more than the cluster merge.
    const size_t ringb = ring_bytes(fmt, rt);
    int S = 0, S_min = 1;
    auto plan = [&](bool xg_) -> int {
        const size_t stepb   = step_bytes(mode, MXs, rows_cta, xg_);
        const int hard_steps = (int)std::max<size_t>(1, (kSmemCap - 256 - ringb) / stepb);
        const int soft_steps = ringb + 4 * stepb <= kSmemSoft ? (int)((kSmemSoft - ringb) / stepb) : 0;
        const int per_sm3 = MXs <= 8 ? KF_GEMV_OCC : (MXs <= 16 ? 2 : 1), per_sm2 = MXs <= 16 ? 2 : 1;
        S_min = (p.steps_total + hard_steps - 1) / hard_steps;
        S     = ctx->gemv_splitk;
        auto waves_of = [&](int cand) {
            const int nst   = (p.steps_total + cand - 1) / cand;
            const int slots = ctx->sm_count * ((soft_steps && nst <= soft_steps) ? per_sm3 : per_sm2);
            return (rb * cand + slots - 1) / slots;
        };
        if (S <= 0) {
            const bool can_cluster = M <= 8 && ctx->gemv_cluster > 0;
            const double bits_of_fmt = (double)fmt_info(fmt).bits;
            double best = 1e30;
            S           = S_min;
            for (int cand = S_min; cand <= std::min(p.steps_total, 64); cand++) {
                const int nst = (p.steps_total + cand - 1) / cand;
                if (nst < 2 && cand > S_min) break;
                // the merge area must fit next to the slice (launch_one falls back to the global merge otherwise)
                // ... and, with several tokens, the slices short: measured at 8 tokens on 5120 x 25600, 8 cluster-merged slices of 25
                // k-steps take 34.6 us against 29.4 for 11 slices through the global workspace
                const bool clustered = can_cluster && cand >= 2 && cand <= 8 && (M == 1 || nst <= 16) &&
                                       ringb + (size_t)nst * stepb + (size_t)cand * M * rows_cta * 4 + 1024 <= kSmemCap;
                if (clustered && (cand & (cand - 1)) && cand > S_min) continue;  // cluster sizes 3, 5, 6, 7 pack badly
                // the cluster merge area ([S][M][rows] floats in the leader) counts against the shared memory that decides 3 or 2 CTAs per SM
                const size_t smem_c = ringb + (size_t)nst * stepb + (clustered ? (size_t)cand * M * rows_cta * 4 : 0);
                const int per_sm  = (soft_steps && nst <= soft_steps && smem_c <= kSmemSoft) ? per_sm3 : per_sm2;
                const int slots   = ctx->sm_count * per_sm;
                const int waves   = (rb * cand + slots - 1) / slots;
                // us: prologue + merge; the merge moves M x rows floats per slice (measured at 8 tokens: global merge ~ +4 us, cluster ~ +1)
                const double fixed = cand == 1 ? 4.0 : clustered ? 5.0 + 0.15 * (MXs - 1) : 7.5 + 0.6 * (MXs - 1);
                // bytes in flight = CTAs x ring depth x 8 KB per k-step; what they can pull per DRAM round trip (~1.65 us) caps the stream
                const double ctas   = std::min<double>((double)rb * cand, slots);
                const double bw     = std::min(6.0e6, ctas * 3.0 * rows_cta * (KSTEP * bits_of_fmt / 8.0) / 1.65);  // bytes per us
                const double stream = std::max(0.55 * nst, (double)total_rows * K * (bits_of_fmt / 8.0 + 4.0 / 128.0) / waves / bw);
                const double cost   = waves * (fixed + stream);
                if (cost < best - 1e-9) best = cost, S = cand;
            }
        }
        return waves_of(std::max(S, S_min));
    };
    // 4..8 tokens of the decode arithmetic: the per-CTA staging area (2 KB per k-step at 8 tokens) caps the k-slice, which on the large
    // shapes forces k-splits that no longer fit one wave.  There the activations are pre-staged ONCE in global memory
    // (kf_gemv_xprep_kernel, one more launch: 51200 x 5120 at 8 tokens 49 -> 34.5 us) and the split is planned without that cap; the
    // small shapes, which fit a wave anyway, keep the per-CTA staging (the extra launch would cost them ~2 us).  Knob gemv_xg_min_m:
    // token count from which this is considered (0 = never; negative = always from |value| tokens, for tests).
    bool xg = false;
    {
        const int thr       = ctx->gemv_xg_min_m;
        const bool eligible = mode == MODE_FAST && rt == 1 && M <= 8 && MXs >= 4 && thr != 0 && M >= (thr < 0 ? -thr : thr);
        const int waves     = plan(false);
        if (eligible && (thr < 0 || waves >= 2)) {
            xg = true;
            plan(true);
            const size_t xbytes = (size_t)p.steps_total * UNITS * MXs * 64, sbytes = (size_t)p.steps_total * MXs * 16;
            int rc = kf_ensure_buf(ctx, &ctx->xg_buf, &ctx->xg_bytes, xbytes + sbytes);
            if (rc) return rc;
            p.xg = p.xg_out = (uint4*)ctx->xg_buf;
            p.sxg = p.sxg_out = (float4*)((uint8_t*)ctx->xg_buf + xbytes);
        }
    }
    S = std::max(S, S_min);
    S = std::max(1, std::min(S, p.steps_total));
    if (M == 1 && ctx->gemv_cluster == 2 && ctx->gemv_splitk <= 0 && S > 8 && S_min <= 8) S = 8;
    p.cluster    = (M <= 8 && ctx->gemv_cluster > 0 && S >= 2 && S <= 8 && (M == 1 || (p.steps_total + S - 1) / S <= 16 || ctx->gemv_splitk > 0)) ? 1 : 0;
    p.S          = S;
    ctx->gemv_last_s = S;
    p.nsteps_max = (p.steps_total + S - 1) / S;  // floor/ceil slicing never exceeds ceil(steps/S)
    if (S > 1 && !p.cluster) {
        int rc = kf_ensure_gemv_ws(ctx, (size_t)S * rb * std::max(8, MXs) * rows_cta * sizeof(float), rb);
        if (rc) return rc;
        p.ws = ctx->gemv_ws, p.cnt = ctx->gemv_cnt;
    }
#define KF_GEMV_CASE(F, MD) \
    if (fmt == F && mode == MD) return launch_nt<F, MD>(ctx, p, rt);
    KF_GEMV_CASE(FMT_Q4, MODE_AFFINE_FMA)
    KF_GEMV_CASE(FMT_Q2, MODE_AFFINE_FMA)
    KF_GEMV_CASE(FMT_Q4, MODE_AFFINE)
    KF_GEMV_CASE(FMT_Q4, MODE_AFFINE_SYM)
    KF_GEMV_CASE(FMT_Q4, MODE_FAST)
    KF_GEMV_CASE(FMT_Q2, MODE_FAST)
    KF_GEMV_CASE(FMT_Q1, MODE_FAST)
    KF_GEMV_CASE(FMT_Q2, MODE_AFFINE)
    KF_GEMV_CASE(FMT_Q2, MODE_AFFINE_SYM)
    KF_GEMV_CASE(FMT_Q2, MODE_SCALE)
    KF_GEMV_CASE(FMT_Q1, MODE_SCALE)
    KF_GEMV_CASE(FMT_F8, MODE_PLAIN)
    KF_GEMV_CASE(FMT_BF16, MODE_PLAIN)
#undef KF_GEMV_CASE
    return KF_ERR_UNSUPPORTED;
}

}  // namespace

int kf_gemv_small(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                  const void* norm_w, float norm_eps) {
#ifdef KF_DEBUG_KNOBS
    if (ctx->debug_skip & 2) return KF_OK;
#endif
    if (M <= 8) {  // the persistent TMA-fed stream-K kernel covers the decode shapes (4-bit, group 128, K % 512 == 0)
        const int rc = kf_gemv_tma(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
        if (rc != 1) return rc;
    }
    return gemv_dispatch(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
}

// ---- tensor-parallel decode: the row-parallel matmul (this rank's K-shard of O / down) whose epilogue IS the exchange (kf_tp.cuh) -----
// y = bf16(residual + bf16(sum over the ranks of x . w^T)), identical on every rank; y may alias residual.  Exchange ordinal = calls since
// kf_tp_begin on this context (every rank issues the same sequence).
extern "C" int kf_linear_exchange(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, const void* residual) {
    if (!ctx || !y || !w || !x || !residual) return KF_ERR_BAD_ARG;
    void* ys[1]  = {y};
    const int rc = gemv_dispatch(ctx, 1, ys, w, x, M, EPI_TP, residual, nullptr, 0.f, ctx->tp_xid);
    if (rc == KF_OK) ctx->tp_xid++;
    return rc;
}
