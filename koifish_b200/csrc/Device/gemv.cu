// gemv.cu -- dequant-fused GEMV / skinny GEMM for decode (M <= 64 tokens):  y[M][N] = x[M][K] . deq(W[N][K])^T
//
// Replaces the reference pair  GTensor::GetDataX (whole-matrix dequant to a bf16 scratch, src/Device/CUDA/kernel/quantizer.cu:
// 249-392, T.cu:245-294)  +  CU_mm_blasLt (cuBLASLt bf16 GEMM, src/Device/CUDA/kernel/gemm.cu:93-214)  as called from
// SLP::Forw (src/Device/CUDA/NeuronFuse.cu:305-381), plus the CU_swiglu_v0 / CU_add3 launches that follow it.
//
// Design (HBM-bound; roofline = bytes of packed weights + gama):
//   * every packed byte is read from HBM exactly once, straight into registers, with 16/8/4-byte loads whose quad-wise union
//     is a contiguous 64/32/16-byte run of one weight row; PF steps are kept in flight per thread;
//   * the codes are expanded in registers to the *bit-exact* bf16 weights the reference's dequant kernel produces
//     (p = RN(step*k), w = RN(p - zero), both bf16) and fed directly as the A fragments of mma.sync.m16n8k16 (bf16 in, fp32
//     accumulate -- the accumulation type cuBLASLt uses in the reference).  Tensor-core *throughput* is irrelevant here (the op
//     is bandwidth bound); the MMA is used because it takes bf16 pairs without an unpack-to-fp32 and does the 16x8x16 FMAs in
//     one issue slot, which is what keeps the CUDA-core instruction count below the HBM rate;
//   * the k-order inside a 128-wide group is permuted to make code extraction cheap (nibbles 16 bits apart form one bf16x2
//     register); the activations are staged in shared memory in the same permuted order, so no weight is ever shuffled;
//   * split-K across CTAs with a deterministic (fixed-order) last-CTA reduction; bias-free epilogues fuse the residual add or
//     SwiGLU(gate, up).
#include <string.h>

#include <algorithm>

#include "kf_common.cuh"

namespace {

enum { FMT_BF16 = 0, FMT_F8 = 1, FMT_Q4 = 2, FMT_Q2 = 3, FMT_Q1 = 4 };
enum { MODE_PLAIN = 0, MODE_AFFINE = 1, MODE_AFFINE_SYM = 2, MODE_SCALE = 3 };
enum { EPI_NONE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_F32 = 4 };

constexpr int kThreads = 256;
constexpr int kWarps   = 8;
constexpr int kRowsCta = 128;
constexpr int kTileStride = 132;  // fp32 tile row stride (padded: conflict-free fragment scatter)

template <int FMT> struct Fmt;
template <> struct Fmt<FMT_Q4>   { static constexpr int KSTEP = 128, UNITS = 4, LOADB = 16, PF = 4, BITS = 4; };
template <> struct Fmt<FMT_Q2>   { static constexpr int KSTEP = 128, UNITS = 4, LOADB = 8,  PF = 6, BITS = 2; };
template <> struct Fmt<FMT_Q1>   { static constexpr int KSTEP = 128, UNITS = 4, LOADB = 4,  PF = 8, BITS = 1; };
template <> struct Fmt<FMT_F8>   { static constexpr int KSTEP = 64,  UNITS = 2, LOADB = 16, PF = 4, BITS = 8; };
template <> struct Fmt<FMT_BF16> { static constexpr int KSTEP = 32,  UNITS = 1, LOADB = 16, PF = 4, BITS = 16; };

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
    int M, K;
    int steps_total;  // K / KSTEP
    int S;            // k-splits
    int total_rb;
    int qbias;
    int gshift;  // log2(group / 128): gama index = row*(K/group) + (step >> gshift)
    int epilogue;
    uint32_t lop_mask, lop_magic;  // code-field mask (0x000F000F / 0x00030003 / 0x00010001) and bf16x2 128.0 (0x43004300)
    float* ws;
    unsigned* cnt;
};

// ---- permuted k-order of the activations inside one thread slot (see header) -------------------------------------------------
template <int FMT>
__host__ __device__ constexpr int xperm(int o) {
    const int u = o >> 3, e = o & 7;
    if (FMT == FMT_Q4) return 8 * u + ((e & 1) ? 3 : 7) - (e >> 1);                               // {7,3,6,2,5,1,4,0}
    if (FMT == FMT_Q2) return 16 * (u >> 1) + ((e & 1) ? 7 : 15) - (e >> 1) - 4 * (u & 1);        // {15,7,14,6,13,5,12,4} / -4
    if (FMT == FMT_Q1) return ((e & 1) ? 15 : 31) - (e >> 1) - 4 * u;                             // {31,15,30,14,29,13,28,12} - 4u
    return o;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int LOADB> struct LoadT;
template <> struct LoadT<16> { using type = uint4; };
template <> struct LoadT<8>  { using type = uint2; };
template <> struct LoadT<4>  { using type = uint32_t; };

__device__ __forceinline__ void ldw(uint4& r, const uint8_t* p) { r = ldg_stream_v4(p); }
__device__ __forceinline__ void ldw(uint2& r, const uint8_t* p) { r = ldg_stream_v2(p); }
__device__ __forceinline__ void ldw(uint32_t& r, const uint8_t* p) { r = ldg_stream_u32(p); }

// 32-bit register #idx (in code order: idx 0 holds the first codes of the thread's slot)
__device__ __forceinline__ uint32_t reg_of(const uint4& q, int idx) { return idx == 0 ? q.w : idx == 1 ? q.z : idx == 2 ? q.y : q.x; }
__device__ __forceinline__ uint32_t reg_of(const uint2& q, int idx) { return idx == 0 ? q.y : q.x; }
__device__ __forceinline__ uint32_t reg_of(const uint32_t& q, int) { return q; }
// natural order (byte / bf16 streams)
__device__ __forceinline__ uint32_t nat_of(const uint4& q, int idx) { return idx == 0 ? q.x : idx == 1 ? q.y : idx == 2 ? q.z : q.w; }

template <int FMT>
struct Stage {
    typename LoadT<Fmt<FMT>::LOADB>::type qa, qb;  // row g, row g+8
    uint32_t ga, gb;                               // zero | step<<16 for the two rows
};

// Expand one pair of codes (16 bits apart in `reg` after the shift) to the bf16x2 weights.
// (a & b) | c in ONE LOP3: mask and magic come from kernel parameters so that ptxas keeps them as register / constant-bank
// operands instead of splitting the op into two immediate-form LOP3s.
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int FMT, int MODE>
__device__ __forceinline__ uint32_t deq_pair(uint32_t reg, int shift, uint32_t step2, uint32_t zero2, uint32_t nbias2, uint32_t bias2,
                                             uint32_t mask, uint32_t magic) {
    uint32_t v = and_or(reg >> shift, mask, magic);  // bf16x2 {128 + c_lo, 128 + c_hi}, exact
    if (MODE == MODE_AFFINE) {
        // qbias == 0:  RN(step*(v-128)) == fma(v, step, -128*step) (single rounding of the exact product step*c) ; then RN(p - zero)
        __nv_bfloat162 p = __hfma2(u32_as_bf162(v), u32_as_bf162(step2), u32_as_bf162(nbias2));
        return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
    } else if (MODE == MODE_AFFINE_SYM) {
        __nv_bfloat162 k = __hsub2(u32_as_bf162(v), u32_as_bf162(bias2));  // exact small integer
        __nv_bfloat162 p = __hmul2_rn(u32_as_bf162(step2), k);
        return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(zero2)));
    } else {  // MODE_SCALE: A = code - qbias ; the group step is applied to the fp32 group sum
        return bf162_as_u32(__hsub2(u32_as_bf162(v), u32_as_bf162(bias2)));
    }
}
// E5M2-by-truncation bytes -> bf16x2 (exact): fp16 bits (b<<8) re-biased into bf16 via a 2^112 multiply
__device__ __forceinline__ uint32_t f8_pair(uint32_t reg, uint32_t sel) {
    uint32_t h2 = __byte_perm(reg, 0u, sel);  // {b_hi<<8 : b_lo<<8} as fp16x2
    uint32_t t  = ((h2 >> 3) & 0x0FE00FE0u) | (h2 & 0x80008000u);
    return bf162_as_u32(__hmul2(u32_as_bf162(t), u32_as_bf162(0x77807780u)));  // * 2^112
}

template <int FMT, int MODE, int NT, bool M1>
__global__ void __launch_bounds__(kThreads, (NT >= 4 ? 1 : 2)) kf_gemv_kernel(const GemvParams p) {
    using F = Fmt<FMT>;
    constexpr int KSTEP = F::KSTEP, UNITS = F::UNITS, PF = F::PF, KT = KSTEP / 4;
    constexpr int MX = M1 ? 1 : 8 * NT;  // token rows staged in shared memory
    constexpr int MP = 8 * NT;           // token columns of the fp32 tile
    extern __shared__ uint4 smem[];
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int rb = blockIdx.x, split = blockIdx.y;
    const bool swiglu = p.epilogue == EPI_SWIGLU;

    // ---- which rows does this warp own? -------------------------------------------------------------------------------------
    int segi = 0;
    if (!swiglu) {
        if (p.nseg > 1 && rb >= p.seg[1].rb0) segi = 1;
        if (p.nseg > 2 && rb >= p.seg[2].rb0) segi = 2;
    } else {
        segi = warp >> 2;
    }
    const GemvSeg& sg = p.seg[segi];
    const int row0    = swiglu ? rb * 64 + (warp & 3) * 16 : (rb - sg.rb0) * kRowsCta + warp * 16;
    const bool active = row0 + 16 <= sg.rows;

    const int s_begin = (int)(((long long)split * p.steps_total) / p.S);
    const int s_end   = (int)(((long long)(split + 1) * p.steps_total) / p.S);
    const int nsteps  = s_end - s_begin;

    // ---- stage the activations of this k-slice in shared memory, permuted to the fragment order --------------------------------
    {
        const int items = MX * nsteps * 4;
        for (int it = tid; it < items; it += kThreads) {
            const int tt = it & 3, m = (it >> 2) % MX, s = it / (4 * MX);
            uint32_t src[KT / 2];
            if (m < p.M) {
                const uint4* gp = reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + (size_t)(s_begin + s) * KSTEP + tt * KT);
#pragma unroll
                for (int i = 0; i < KT / 8; i++) {
                    uint4 v = __ldg(gp + i);
                    src[4 * i + 0] = v.x, src[4 * i + 1] = v.y, src[4 * i + 2] = v.z, src[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < KT / 2; i++) src[i] = 0u;
            }
#pragma unroll
            for (int u = 0; u < UNITS; u++) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int e0 = xperm<FMT>(u * 8 + 2 * j), e1 = xperm<FMT>(u * 8 + 2 * j + 1);
                    const uint32_t lo = (src[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu;
                    const uint32_t hi = (src[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu;
                    o[j] = lo | (hi << 16);
                }
                smem[((s * UNITS + u) * MX + m) * 4 + tt] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    __syncthreads();

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[nt][j] = 0.f;

    if (active && nsteps > 0) {
        const size_t row_bytes = (size_t)p.K * F::BITS / 8;
        int toff;  // byte offset of this thread's slot inside one k-step of a row
        if (FMT == FMT_Q2)
            toff = 16 * (t >> 1) + 8 * (1 - (t & 1));  // word.high holds the first 32 codes (PackedQ.hpp:185-198)
        else if (FMT == FMT_Q1)
            toff = 12 - 4 * t;                          // high.hi32 holds codes 0..31 (PackedQ.hpp:200-211)
        else
            toff = 16 * t;
        constexpr int STEPB = KSTEP * F::BITS / 8;  // bytes per row per k-step
        const uint8_t* pa = sg.data + (size_t)(row0 + g) * row_bytes + (size_t)s_begin * STEPB + toff;
        const uint8_t* pb = pa + 8 * row_bytes;
        const int gpr     = (p.K >> 7) >> p.gshift;  // groups per row
        const uint16_t *za = nullptr, *sa = nullptr, *zb = nullptr, *sb = nullptr;
        if (MODE != MODE_PLAIN) {
            za = sg.zero + (size_t)(row0 + g) * gpr, sa = sg.step + (size_t)(row0 + g) * gpr;
            zb = za + (size_t)8 * gpr, sb = sa + (size_t)8 * gpr;
        }
        const uint32_t bias2 = pack_bf16x2((float)(128 + p.qbias), (float)(128 + p.qbias));

        Stage<FMT> st[PF];
        auto load_stage = [&](Stage<FMT>& s_, int sl) {
            ldw(s_.qa, pa + (size_t)sl * STEPB);
            ldw(s_.qb, pb + (size_t)sl * STEPB);
            if (MODE != MODE_PLAIN) {
                const int gi = (s_begin + sl) >> p.gshift;
                s_.ga = (uint32_t)__ldg(za + gi) | ((uint32_t)__ldg(sa + gi) << 16);
                s_.gb = (uint32_t)__ldg(zb + gi) | ((uint32_t)__ldg(sb + gi) << 16);
            }
        };
#pragma unroll
        for (int i = 0; i < PF; i++)
            if (i < nsteps) load_stage(st[i], i);

        const int xlane = (M1 ? 0 : g) * 4 + t;
        for (int s0 = 0; s0 < nsteps; s0 += PF) {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const int s = s0 + i;
                if (s >= nsteps) break;
                const Stage<FMT>& cur = st[i];

                uint32_t step2a = 0, zero2a = 0, nb2a = 0, step2b = 0, zero2b = 0, nb2b = 0;
                if (MODE != MODE_PLAIN) {
                    step2a = __byte_perm(cur.ga, 0u, 0x3232), zero2a = __byte_perm(cur.ga, 0u, 0x1010);
                    step2b = __byte_perm(cur.gb, 0u, 0x3232), zero2b = __byte_perm(cur.gb, 0u, 0x1010);
                    if (MODE == MODE_AFFINE) {
                        nb2a = bf162_as_u32(__hmul2(u32_as_bf162(step2a), u32_as_bf162(0xC300C300u)));  // -128*step, exact
                        nb2b = bf162_as_u32(__hmul2(u32_as_bf162(step2b), u32_as_bf162(0xC300C300u)));
                    }
                }
                float accg[NT][4];
                if (MODE == MODE_SCALE) {
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int j = 0; j < 4; j++) accg[nt][j] = 0.f;
                }
#pragma unroll
                for (int u = 0; u < UNITS; u++) {
                    uint4 xb[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) xb[nt] = smem[((s * UNITS + u) * MX + (M1 ? 0 : nt * 8)) * 4 + xlane];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t a[4];
                        if constexpr (FMT == FMT_Q4) {
                            const uint32_t ra = reg_of(cur.qa, u), rb_ = reg_of(cur.qb, u);
                            a[0] = deq_pair<FMT, MODE>(ra, 8 * h, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[1] = deq_pair<FMT, MODE>(rb_, 8 * h, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                            a[2] = deq_pair<FMT, MODE>(ra, 8 * h + 4, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[3] = deq_pair<FMT, MODE>(rb_, 8 * h + 4, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                        } else if constexpr (FMT == FMT_Q2) {
                            const uint32_t ra = reg_of(cur.qa, u >> 1), rb_ = reg_of(cur.qb, u >> 1);
                            const int m4 = 4 * (2 * (u & 1) + h);
                            a[0] = deq_pair<FMT, MODE>(ra, m4, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[1] = deq_pair<FMT, MODE>(rb_, m4, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                            a[2] = deq_pair<FMT, MODE>(ra, m4 + 2, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[3] = deq_pair<FMT, MODE>(rb_, m4 + 2, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                        } else if constexpr (FMT == FMT_Q1) {
                            const uint32_t ra = reg_of(cur.qa, 0), rb_ = reg_of(cur.qb, 0);
                            const int m2 = 2 * (2 * u + h);
                            a[0] = deq_pair<FMT, MODE>(ra, m2, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[1] = deq_pair<FMT, MODE>(rb_, m2, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                            a[2] = deq_pair<FMT, MODE>(ra, m2 + 1, step2a, zero2a, nb2a, bias2, p.lop_mask, p.lop_magic);
                            a[3] = deq_pair<FMT, MODE>(rb_, m2 + 1, step2b, zero2b, nb2b, bias2, p.lop_mask, p.lop_magic);
                        } else if constexpr (FMT == FMT_F8) {
                            const uint32_t ra = nat_of(*reinterpret_cast<const uint4*>(&cur.qa), 2 * u + h);
                            const uint32_t rb_ = nat_of(*reinterpret_cast<const uint4*>(&cur.qb), 2 * u + h);
                            a[0] = f8_pair(ra, 0x1404u), a[1] = f8_pair(rb_, 0x1404u);
                            a[2] = f8_pair(ra, 0x3424u), a[3] = f8_pair(rb_, 0x3424u);
                        } else {  // FMT_BF16
                            a[0] = nat_of(*reinterpret_cast<const uint4*>(&cur.qa), 2 * h);
                            a[1] = nat_of(*reinterpret_cast<const uint4*>(&cur.qb), 2 * h);
                            a[2] = nat_of(*reinterpret_cast<const uint4*>(&cur.qa), 2 * h + 1);
                            a[3] = nat_of(*reinterpret_cast<const uint4*>(&cur.qb), 2 * h + 1);
                        }
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            const uint32_t b0 = h ? xb[nt].z : xb[nt].x, b1 = h ? xb[nt].w : xb[nt].y;
                            if (MODE == MODE_SCALE)
                                mma_bf16_16816(accg[nt], a, b0, b1);
                            else
                                mma_bf16_16816(acc[nt], a, b0, b1);
                        }
                    }
                }
                if (MODE == MODE_SCALE) {
                    const float fa = bf16hi(cur.ga), fb = bf16hi(cur.gb);  // step of row g / row g+8
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        acc[nt][0] = fmaf(fa, accg[nt][0], acc[nt][0]);
                        acc[nt][1] = fmaf(fa, accg[nt][1], acc[nt][1]);
                        acc[nt][2] = fmaf(fb, accg[nt][2], acc[nt][2]);
                        acc[nt][3] = fmaf(fb, accg[nt][3], acc[nt][3]);
                    }
                }
                if (s + PF < nsteps) load_stage(st[i], s + PF);  // slot is dead now: refill it PF steps ahead
            }
        }
    }

    // ---- scatter fragments to the fp32 tile [MP][128 rows] in shared memory -------------------------------------------------
    __syncthreads();  // everyone is done reading x
    float* tile = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const int m0 = nt * 8 + 2 * t, r = warp * 16 + g;
        tile[(m0 + 0) * kTileStride + r]     = acc[nt][0];
        tile[(m0 + 1) * kTileStride + r]     = acc[nt][1];
        tile[(m0 + 0) * kTileStride + r + 8] = acc[nt][2];
        tile[(m0 + 1) * kTileStride + r + 8] = acc[nt][3];
    }
    __syncthreads();

    // ---- split-K: publish the partial tile; the last CTA of this row block reduces in fixed order ---------------------------
    if (p.S > 1) {
        float* wsp = p.ws + ((size_t)split * p.total_rb + rb) * (size_t)(MP * kRowsCta);
        for (int e = tid; e < p.M * kRowsCta; e += kThreads) {
            const int m = e >> 7, r = e & 127;
            __stcg(wsp + m * kRowsCta + r, tile[m * kTileStride + r]);
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
        for (int e = tid; e < p.M * kRowsCta; e += kThreads) {
            const int m = e >> 7, r = e & 127;
            float sum = 0.f;
            for (int sp = 0; sp < p.S; sp++) sum += __ldcg(p.ws + ((size_t)sp * p.total_rb + rb) * (size_t)(MP * kRowsCta) + m * kRowsCta + r);
            tile[m * kTileStride + r] = sum;
        }
        if (tid == 0) p.cnt[rb] = 0u;  // self-reset for the next launch
        __syncthreads();
    }

    // ---- epilogue -----------------------------------------------------------------------------------------------------------
    if (!swiglu) {
        const int rbase = (rb - sg.rb0) * kRowsCta;
        for (int e = tid; e < p.M * kRowsCta; e += kThreads) {
            const int m = e >> 7, r = e & 127, row = rbase + r;
            if (row >= sg.rows) continue;
            if (p.epilogue == EPI_F32) {  // tensor-parallel partial sums stay fp32 until the all-reduce
                reinterpret_cast<float*>(sg.y)[(size_t)m * sg.rows + row] = tile[m * kTileStride + r];
                continue;
            }
            uint16_t v = f32_to_bf16_bits(tile[m * kTileStride + r]);  // the reference's GEMM writes bf16 (gemm.cu:124-126)
            if (p.epilogue == EPI_RESIDUAL)                            // then CU_add3 adds the residual in fp32 (packedN.cuh:867-875)
                v = f32_to_bf16_bits(bf16_bits_to_f32(p.residual[(size_t)m * sg.rows + row]) + bf16_bits_to_f32(v));
            sg.y[(size_t)m * sg.rows + row] = v;
        }
    } else {
        const int rows = p.seg[0].rows;
        for (int e = tid; e < p.M * 64; e += kThreads) {
            const int m = e >> 6, r = e & 63, row = rb * 64 + r;
            if (row >= rows) continue;
            const float gt = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * kTileStride + r]));
            const float up = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * kTileStride + 64 + r]));
            p.seg[0].y[(size_t)m * rows + row] = f32_to_bf16_bits((gt * up) / (1.0f + expf(-gt)));  // CU_swiglu_v0, Activation.cu:86-93
        }
    }
}

template <int FMT, int MODE, int NT, bool M1>
int launch_one(kf_ctx* ctx, const GemvParams& p, int nsteps_max) {
    using F = Fmt<FMT>;
    constexpr int MX = M1 ? 1 : 8 * NT, MP = 8 * NT;
    size_t xbytes    = (size_t)nsteps_max * F::UNITS * MX * 4 * 16;
    size_t tilebytes = (size_t)MP * kTileStride * 4;
    size_t smem      = std::max(xbytes, tilebytes);
    auto kern        = kf_gemv_kernel<FMT, MODE, NT, M1>;
    static size_t smem_set = 0;  // per instantiation
    if (smem > 48 * 1024 && smem > smem_set) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        smem_set = 100 * 1024;
    }
    dim3 grid(p.total_rb, p.S);
    kern<<<grid, kThreads, smem, ctx->stream>>>(p);
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}

template <int FMT, int MODE>
int launch_nt(kf_ctx* ctx, const GemvParams& p, int nsteps_max) {
    if (p.M == 1) return launch_one<FMT, MODE, 1, true>(ctx, p, nsteps_max);
    if (p.M <= 8) return launch_one<FMT, MODE, 1, false>(ctx, p, nsteps_max);
    if (p.M <= 16) return launch_one<FMT, MODE, 2, false>(ctx, p, nsteps_max);
    if (p.M <= 32) return launch_one<FMT, MODE, 4, false>(ctx, p, nsteps_max);
    return launch_one<FMT, MODE, 8, false>(ctx, p, nsteps_max);
}

int gemv_dispatch(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    KF_REQUIRE(ctx, n >= 1 && n <= 3 && M >= 1 && M <= 64 && x, "1..3 weights, 1..64 tokens");
    const int type = w[0].type, K = w[0].cols;
    int fmt, mode;
    switch (type) {
        case KF_T_BF16: fmt = FMT_BF16, mode = MODE_PLAIN; break;
        case KF_T_F8E5M2: fmt = FMT_F8, mode = MODE_PLAIN; break;
        case KF_T_Q4: fmt = FMT_Q4, mode = w[0].qbias == 0 ? MODE_AFFINE : MODE_AFFINE_SYM; break;
        case KF_T_Q2: fmt = FMT_Q2, mode = w[0].qbias == 0 ? MODE_AFFINE : MODE_AFFINE_SYM; break;
        case KF_T_SIGN: fmt = FMT_Q2, mode = MODE_SCALE; break;
        case KF_T_BINARY: fmt = FMT_Q1, mode = MODE_SCALE; break;
        default: return KF_ERR_UNSUPPORTED;
    }
    const int kstep = fmt == FMT_BF16 ? 32 : fmt == FMT_F8 ? 64 : 128;
    KF_REQUIRE(ctx, K % kstep == 0 && K % 8 == 0, "K must be a multiple of the k-step");
    GemvParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = n, p.x = (const uint16_t*)x, p.residual = (const uint16_t*)residual, p.M = M, p.K = K;
    p.steps_total = K / kstep, p.qbias = w[0].qbias, p.epilogue = epilogue;
    p.lop_mask = fmt == FMT_Q4 ? 0x000F000Fu : fmt == FMT_Q2 ? 0x00030003u : 0x00010001u, p.lop_magic = 0x43004300u;
    int rb = 0;
    for (int i = 0; i < n; i++) {
        KF_REQUIRE(ctx, w[i].type == type && w[i].cols == K && w[i].qbias == w[0].qbias && w[i].group == w[0].group,
                   "fused weights must share type / K / quant card");
        KF_REQUIRE(ctx, w[i].rows % 16 == 0 && w[i].data_dev && y[i], "rows must be a multiple of 16");
        p.seg[i].data = (const uint8_t*)w[i].data_dev, p.seg[i].y = (uint16_t*)y[i], p.seg[i].rows = w[i].rows, p.seg[i].rb0 = rb;
        if (mode != MODE_PLAIN) {
            KF_REQUIRE(ctx, kf_has_gama(w[i]) && w[i].group >= 128 && (w[i].group & (w[i].group - 1)) == 0 && K % w[i].group == 0,
                       "fused path needs group = 128 * 2^n dividing K");
            p.seg[i].zero = kf_gama_zero(w[i]), p.seg[i].step = kf_gama_step(w[i]);
        }
        rb += (w[i].rows + kRowsCta - 1) / kRowsCta;
    }
    if (mode != MODE_PLAIN) {
        int gs = 0;
        while ((128 << gs) < w[0].group) gs++;
        p.gshift = gs;
    }
    if (epilogue == EPI_SWIGLU) {
        KF_REQUIRE(ctx, n == 2 && w[0].rows == w[1].rows, "swiglu needs gate and up of equal shape");
        rb = (w[0].rows + 63) / 64;
    }
    if (epilogue == EPI_RESIDUAL) KF_REQUIRE(ctx, n == 1 && residual, "residual epilogue takes one weight");
    p.total_rb = rb;

    // ---- k-split heuristic: enough CTAs for ~2 waves of 2 resident CTAs per SM, slices that fit shared memory ----------------
    const int MXs        = M == 1 ? 1 : (M <= 8 ? 8 : M <= 16 ? 16 : M <= 32 ? 32 : 64);
    const int units      = fmt == FMT_BF16 ? 1 : fmt == FMT_F8 ? 2 : 4;
    const size_t stepsm  = (size_t)units * MXs * 64;  // shared bytes per k-step
    const int max_steps  = (int)std::max<size_t>(1, (96 * 1024) / stepsm);
    int S                = ctx->gemv_splitk;
    if (S <= 0) {
        const int target = ctx->sm_count * 4;
        S                = (target + rb - 1) / rb;
        const int min_steps = fmt == FMT_BF16 ? 16 : 8;
        S                = std::min(S, std::max(1, p.steps_total / min_steps));
        S                = std::min(S, 32);
    }
    S = std::max(S, (p.steps_total + max_steps - 1) / max_steps);
    S = std::max(1, std::min(S, p.steps_total));
    p.S = S;
    const int nsteps_max = (p.steps_total + S - 1) / S;  // floor/ceil slicing never exceeds ceil(steps/S) <= max_steps
    if (S > 1) {
        int rc = kf_ensure_gemv_ws(ctx, (size_t)S * rb * MXs * kRowsCta * sizeof(float) * (M == 1 ? 8 : 1), rb);
        if (rc) return rc;
        p.ws = ctx->gemv_ws, p.cnt = ctx->gemv_cnt;
    }
#define KF_GEMV_CASE(F, MD) \
    if (fmt == F && mode == MD) return launch_nt<F, MD>(ctx, p, nsteps_max);
    KF_GEMV_CASE(FMT_Q4, MODE_AFFINE)
    KF_GEMV_CASE(FMT_Q4, MODE_AFFINE_SYM)
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

int kf_gemv_small(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    return gemv_dispatch(ctx, n, y, w, x, M, epilogue, residual);
}
