// gemv_tma.cu -- persistent, TMA-fed, stream-K dequant-GEMV for decode (1..8 tokens) over 4-bit PackedQ weights:
//     y[M][N] = x[M][K] . deq(W[N][K])^T      (+ fused RMSNorm of x, + residual / SwiGLU / fp32-partial epilogues)
//
// Replaces the reference pair GTensor::GetDataX (whole-matrix dequant to a bf16 scratch, src/Device/CUDA/kernel/quantizer.cu:249-392,
// CU_Q128toX_ T.cu:245-294) + CU_mm_blasLt (cuBLASLt GEMM, src/Device/CUDA/kernel/gemm.cu:93-214) as called from SLP::Forw
// (src/Device/CUDA/NeuronFuse.cu:305-381), with the neighbouring CU_rms_infer / CU_swiglu_v0 / CU_add3 launches folded in.
//
// The op is HBM-bound on paper (0.53 B / weight, every packed byte read once).  What actually limited round 1's kernel (gemv.cu) was the
// ALU pipe: tools/ubench/deq_rate*.cu measure, from registers only, 432 clk per [128 rows x 128 k] tile and SM for the bit-exact bf16
// unpack (SHF + LOP3 + HSUB2 + HFMA2 per pair of weights; a funnel shift alone costs ~5 clk per warp instruction and scheduler on this
// part) -- MORE than the 392 clk the tile's bytes take at the measured HBM rate.  Hence two arithmetic modes (template MODE):
//   DM_FAST (default)  the codes go to the tensor cores as fp16 numbers 1024 + c / 1024 + 16 c, built with ONE LOP3 each (both nibbles of
//                      a byte sit inside fp16's 10 mantissa bits; one byte-permute per 8 codes replaces six funnel shifts): ~105-160 clk
//                      per tile.  The group's step / zero / code bias are applied to the fp32 group sums:
//                          y += step * (sum_k c_k x_k - qbias * Sx) - zero * Sx,   Sx = sum_k x_k
//                      i.e. the reference's affine dequant WITHOUT its intermediate bf16 rounding of every weight (logits agree within the
//                      stated tolerance; the dequantised weights themselves stay bit-exact through kf_dequant).  Activations are staged as
//                      fp16 with a power-of-two scale per 128-k group (exact for bf16 inputs over 28 binades).
//   DM_FMA / DM_TWO    the reference's expression evaluated per weight in bf16 with one / two roundings (ctx gemv_exact = 1 with
//                      deq_fma = 1 / 0): bit-exact weights inside the matmul, ALU-bound at ~0.6 of the HBM peak.
// Structure (all modes):
//   * PERSISTENT stream-K grid: one CTA per SM (or two: gemv_tma_occ) walks an equal, contiguous share of the (row block, k step) tiles:
//     the prologue (RMSNorm sum, activation staging) is paid once per SM and all SMs finish together whatever the shape; a row block
//     that spans CTAs is reduced in a fixed order (partials through L2, the last contributor adds them in CTA order: bit-reproducible);
//   * ONE producer thread per CTA streams [128 rows x 4 k steps] stages of packed bytes with TMA (cp.async.bulk.tensor, UTMALDG) into an
//     mbarrier ring, and the zero / step tiles of 8 k steps into a second, small ring; the consumer warps only do LDS + unpack +
//     mma.sync: no per-thread address arithmetic, no cp.async bookkeeping;
//   * the shared-memory / register footprint of the 8-warp variant is below half an SM, so the NEXT kernel of the stream (programmatic
//     dependent launch) becomes resident and fills its own ring while this one drains: the HBM stream does not stop at kernel boundaries.
#include <string.h>

#include <algorithm>
#include <map>

#include "kf_async.cuh"

namespace {
using namespace kfa;

enum { DM_FMA = 0, DM_TWO = 1, DM_FAST = 2 };
enum { EPI_NONE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_F32 = 4 };

constexpr int UROWS     = 128;             // weight rows per tile (8 warps x 16 rows)
constexpr int UBYTES    = UROWS * 64;      // one k step of a tile: 4-bit, 64 bytes per row
constexpr int CHUNK     = 4;               // k steps per ring stage
constexpr int SBYTES    = CHUNK * UBYTES;  // 32 KB per stage
constexpr int TS        = UROWS + 4;       // row stride of the fp32 result tile
constexpr int kMaxStage = 8;
constexpr int GBLK      = 8;                // k steps (= quantisation groups) covered by one zero / step tile
constexpr int GBYTES    = 2 * UROWS * GBLK * 2;  // [zero | step] x [128 rows][8 groups] bf16 = 4 KB
constexpr int kGStage   = 4;                // depth of the zero / step ring (4 x 8 k steps >= the weight ring + one block)
struct alignas(64) TMaps {
    CUtensorMap w[3];  // packed codes  [rows][K/2 bytes], box 64 rows x 64 bytes
    CUtensorMap z[3];  // zero          [rows][K/128] bf16, box 64 rows x 8 groups
    CUtensorMap s[3];  // step
};

struct TSeg {
    const uint16_t* zero;
    const uint16_t* step;
    uint16_t* y;
    int rows;
    int rb0;  // first row block of this weight in the launch's row-block space
};
struct TParams {
    TSeg seg[3];
    int nseg;
    const uint16_t* x;
    const uint16_t* residual;
    const uint16_t* norm_w;
    float norm_eps;
    int M, K;
    int ksteps;    // K / 128
    int total_rb;  // 128-row blocks (SwiGLU: 64 gate rows + 64 up rows)
    int units;     // total_rb * ksteps
    int upc;       // k-step tiles per CTA = ceil(units / grid)
    int nstage;    // ring depth
    int xs;        // k steps of activations staged at once (multiple of 4)
    int qbias, epilogue;
    int dbg;  // -DKF_DEBUG_KNOBS builds only (timing experiments, results are garbage): 1 = consumers skip the math, 2 = no weight TMA
    float* ws;        // [grid][8][UROWS] partial tiles of row blocks that span CTAs
    unsigned* flags;  // [grid] "partial of CTA c is in ws" (self-resetting)
    // byte offsets inside dynamic shared memory
    int off_g, off_x, off_sx, off_tile, off_stash, off_bar;
};

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
// (a & IMM) | c in ONE LOP3 with the mask as an immediate: two register reads (no bank conflict on the third source)
template <uint32_t IMM>
__device__ __forceinline__ uint32_t and_or_imm(uint32_t a, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c));
    return d;
}
// permuted k order inside a 32-weight slot (gemv.cu xperm<FMT_Q4>): codes 16 bits apart in a register form one 16-bit pair
__host__ __device__ constexpr int xperm4(int o) { return 8 * (o >> 3) + ((o & 1) ? 3 : 7) - ((o & 7) >> 1); }

// bit-exact modes: one pair of codes (16 bits apart in `reg` after the shift) -> bf16x2 weights.  gz = -zero (DM_FMA) / zero (DM_TWO)
template <int MODE>
__device__ __forceinline__ uint32_t deq_pair(uint32_t reg, int shift, uint32_t step2, uint32_t gz, uint32_t bias2, uint32_t magic) {
    const uint32_t v = and_or_imm<0x000F000Fu>(reg >> shift, magic);            // bf16x2 {128 + c_lo, 128 + c_hi}, exact
    const __nv_bfloat162 k = __hsub2_rn(u32_as_bf162(v), u32_as_bf162(bias2));  // code - qbias: a small integer, exact
    if (MODE == DM_FMA) return bf162_as_u32(__hfma2(k, u32_as_bf162(step2), u32_as_bf162(gz)));  // RN(step*k - zero): ONE rounding (T.cu:274 as built for sm_90+)
    const __nv_bfloat162 p = __hmul2_rn(u32_as_bf162(step2), k);                                   // RN(step*k)
    return bf162_as_u32(__hsub2_rn(p, u32_as_bf162(gz)));                                          // RN(p - zero): TWO roundings
}

// MX: token rows staged in shared memory (1, 2, 4 or 8; the MMA's 8 columns replicate them) ; NWG: consumer warp groups of 8 warps
// (group w takes k steps [w * 4 / NWG, (w + 1) * 4 / NWG) of every stage)
template <int MODE, int MX, int NWG>
__global__ void __launch_bounds__(256 * NWG + 32, NWG == 1 ? 2 : 1)
kf_gemv_tma_kernel(const __grid_constant__ TMaps tm, const TParams p) {
    constexpr int kCW = 8 * NWG, kCT = 256 * NWG, UPW = CHUNK / NWG;  // consumer warps / threads, k steps per warp group and stage
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ float s_red[8];
    __shared__ float s_scale[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    kf_grid_launch_dependents();  // the next kernel of the stream may become resident and start its own weight stream

    const int u0 = blockIdx.x * p.upc, u1 = min(u0 + p.upc, p.units), n = u1 - u0;
    const int NS = p.nstage;
    const uint32_t ring0 = smem_u32(smem);
    const uint32_t gring0 = smem_u32(smem + p.off_g);
    const uint32_t full0 = smem_u32(smem + p.off_bar), empty0 = full0 + 8 * kMaxStage;
    const uint32_t gfull0 = empty0 + 8 * kMaxStage, gempty0 = gfull0 + 8 * kGStage;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(full0 + 8 * s, 1), mbar_init(empty0 + 8 * s, kCW);
        for (int s = 0; s < kGStage; s++) mbar_init(gfull0 + 8 * s, 1), mbar_init(gempty0 + 8 * s, kCW);
        mbar_init_fence();
    }
    __syncthreads();
    if (n <= 0) return;
    const bool swiglu = p.epilogue == EPI_SWIGLU;
    const int ubase = u0 & ~(CHUNK - 1);  // stages hold the k steps [uc, uc + 4) /\ [u0, u1) for uc = ubase, ubase + 4, ...

    // =================================================================================================== producer warp ============
    // One thread feeds both rings.  It runs serially (every instruction waits for the previous one), so the loop is kept lean: one
    // barrier round trip per 4-step stage, the per-row-block quantities recomputed only when the row block changes.
    if (warp == kCW) {
        if (lane == 0) {
            int s = 0, ph = 0, gs = 0, gph = 0, gcount = 0, scount = 0;
            int rb = ubase / p.ksteps, ks0 = ubase - rb * p.ksteps;
            int cur_rb = -1, ra = 0, rbw = 0;
            const CUtensorMap *wa = nullptr, *wb = nullptr, *za = nullptr, *zb = nullptr, *sa = nullptr, *sb = nullptr;
            for (int uc = ubase; uc < u1; uc += CHUNK) {
                if (rb != cur_rb) {  // weight (segment) and first rows of the two 64-row halves of this row block
                    cur_rb = rb;
                    int si = 0, sj;
                    if (swiglu) {
                        si = 0, sj = 1, ra = rbw = rb * 64;
                    } else {
                        if (p.nseg > 1 && rb >= p.seg[1].rb0) si = 1;
                        if (p.nseg > 2 && rb >= p.seg[2].rb0) si = 2;
                        sj = si, ra = (rb - p.seg[si].rb0) * UROWS, rbw = ra + 64;
                    }
                    wa = &tm.w[si], wb = &tm.w[sj], za = &tm.z[si], zb = &tm.z[sj], sa = &tm.s[si], sb = &tm.s[sj];
                }
                if (uc == ubase || (ks0 & (GBLK - 1)) == 0) {  // a new block of 8 groups: its zero / step tiles go first
                    if (gcount >= kGStage) mbar_wait(gempty0 + 8 * gs, gph ^ 1);
                    const uint32_t dst = gring0 + gs * GBYTES, bar = gfull0 + 8 * gs;
                    const int g0 = ks0 & ~(GBLK - 1);
                    mbar_arrive_expect_tx(bar, GBYTES);
                    tma_load_2d(dst, za, g0, ra, bar);
                    tma_load_2d(dst + GBYTES / 4, zb, g0, rbw, bar);
                    tma_load_2d(dst + GBYTES / 2, sa, g0, ra, bar);
                    tma_load_2d(dst + 3 * GBYTES / 4, sb, g0, rbw, bar);
                    gcount++;
                    if (++gs == kGStage) gs = 0, gph ^= 1;
                }
                const int jlo = max(u0 - uc, 0), jhi = min(CHUNK, u1 - uc);
                if (scount >= NS) mbar_wait(empty0 + 8 * s, ph ^ 1);  // the consumers have read the previous contents of this stage
                const uint32_t dst = ring0 + s * SBYTES, bar = full0 + 8 * s;
#ifdef KF_DEBUG_KNOBS
                if (p.dbg & 2) {
                    mbar_arrive(bar);
                } else
#endif
                {
                    mbar_arrive_expect_tx(bar, UBYTES * (jhi - jlo));  // rows beyond the tensor are zero-filled and still counted
                    for (int j = jlo; j < jhi; j++) {
                        tma_load_2d(dst + j * UBYTES, wa, (ks0 + j) * 64, ra, bar);
                        tma_load_2d(dst + j * UBYTES + UBYTES / 2, wb, (ks0 + j) * 64, rbw, bar);
                    }
                }
                scount++;
                if (++s == NS) s = 0, ph ^= 1;
                ks0 += CHUNK;
                if (ks0 == p.ksteps) ks0 = 0, rb++;
            }
        }
        return;
    }

    // =================================================================================================== consumer warps ===========
    const int g = lane >> 2, t = lane & 3;
    const int wg = warp >> 3, wl = warp & 7;  // warp group (which k steps of a stage) / warp inside the group (which 16 rows)
    auto cbar = [] { named_bar_sync(1, kCT); };
    uint4* xs    = reinterpret_cast<uint4*>(smem + p.off_x);      // [slot][4 units][MX][4 thread slots] x 16 bytes
    float4* sxs  = reinterpret_cast<float4*>(smem + p.off_sx);    // DM_FAST: [slot][MX] {1024 Sxe + qbias Sx, Sx, 1 / scale, -} of the staged group
    float* tile  = reinterpret_cast<float*>(smem + p.off_tile);   // [NWG][MX][TS]: one partial tile per warp group
    float* stash = reinterpret_cast<float*>(smem + p.off_stash);  // [MX][TS] partial of the row block this CTA will reduce at the end
    constexpr int TILE1 = MX * TS;

    // rows of this thread inside a tile: warp wl of a group owns rows 16 wl .. 16 wl + 15 (halves: warps 0-3 / 4-7), thread rows g, g + 8
    const int trow = 16 * wl + g;
    const uint32_t woff = (uint32_t)(UPW * wg) * UBYTES + (uint32_t)trow * 64 + 16 * t;

    // ---- everything above touched nothing the previous kernel writes.  From here on we read x / residual and write y / ws ----------
    kf_grid_dependency_wait();

    // ---- optional fused RMSNorm: the arithmetic and summation order of kf_rmsnorm_kernel (ops.cu: 256 threads), bit for bit ----------
    if (p.norm_w) {
        for (int m = 0; m < p.M; m++) {
            if (tid < 256) {
                const uint16_t* xr = p.x + (size_t)m * p.K;
                float ss = 0.f;
                for (int i = tid * 8; i < p.K; i += 256 * 8) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xr + i));
                    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float a = bf16lo(q[j]), b = bf16hi(q[j]);
                        ss = fmaf(a, a, ss), ss = fmaf(b, b, ss);
                    }
                }
                ss = warp_sum(ss);
                if (lane == 0) s_red[warp] = ss;
            }
            cbar();
            if (tid == 0) {
                float tot = 0.f;
                for (int i = 0; i < 8; i++) tot += s_red[i];
                s_scale[m] = 1.0f / sqrtf(fmaf(tot, 1.0f / (float)p.K, p.norm_eps));
            }
            cbar();
        }
    }

    float acc[2][4];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[c][j] = 0.f;
    const uint32_t bias2 = pack_bf16x2((float)(128 + p.qbias), (float)(128 + p.qbias));
    const uint32_t magic = MODE == DM_FAST ? 0x64006400u : 0x43004300u;  // fp16x2 1024.0 / bf16x2 128.0
    const int xcol       = (MX >= 8 ? g : (g & (MX - 1))) * 4 + t;       // column g of the MMA reads token g mod MX
    int stage = 0, phase = 0, gstage = 0, gphase = 0;
    int pend_rb = -1;             // row block whose reduction this CTA owns (its first, partial, segment), done after the last tile
    int seg_k0  = u0 % p.ksteps;  // first k step of the segment being accumulated

    // ---- epilogue of one finished row block whose full sums sit in `tile` (group 0's slot) ---------------------------------------------
    auto epilogue = [&](int rb) {
        if (!swiglu) {
            int si = 0;
            if (p.nseg > 1 && rb >= p.seg[1].rb0) si = 1;
            if (p.nseg > 2 && rb >= p.seg[2].rb0) si = 2;
            const TSeg& sg  = p.seg[si];
            const int rbase = (rb - sg.rb0) * UROWS;
            for (int e = tid; e < p.M * UROWS; e += kCT) {
                const int m = e / UROWS, r = e % UROWS, row = rbase + r;
                if (row >= sg.rows) continue;
                if (p.epilogue == EPI_F32) {  // tensor-parallel partial sums stay fp32 until the exchange
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
            for (int e = tid; e < p.M * 64; e += kCT) {
                const int m = e / 64, r = e % 64, row = rb * 64 + r;
                if (row >= rows) continue;
                const float gt = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * TS + r]));
                const float up = bf16_bits_to_f32(f32_to_bf16_bits(tile[m * TS + 64 + r]));
                p.seg[0].y[(size_t)m * rows + row] = f32_to_bf16_bits((gt * up) / (1.0f + expf(-gt)));  // CU_swiglu_v0, Activation.cu:86-93
            }
        }
    };

    // ---- a segment (row block rb, k steps [k0, k1)) of this CTA is complete: full result, or a partial for the stream-K reduction ----
    auto flush = [&](int rb, int k0, int k1) {
        // fragments -> tile[group][m][row]; thread (g, t) holds columns 2t, 2t+1 of rows g / g+8 of its warp
        float* mine = tile + wg * TILE1;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int col = 2 * t + c;
            if (col < MX && col < p.M) {
                mine[col * TS + trow]     = acc[0][c] + acc[1][c];
                mine[col * TS + trow + 8] = acc[0][2 + c] + acc[1][2 + c];
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[c][j] = 0.f;
        cbar();
        if (NWG > 1) {  // the warp groups hold different k steps of the same rows: add them in a fixed order into group 0's slot
            for (int e = tid; e < p.M * UROWS; e += kCT) {
                const int i = (e / UROWS) * TS + (e % UROWS);
                tile[i] = tile[i] + tile[TILE1 + i];
            }
            cbar();
        }
        if (k0 == 0 && k1 == p.ksteps) {
            epilogue(rb);
        } else if (k1 == p.ksteps) {  // we hold the LAST k steps of rb: we are its reducer; keep our share, reduce after the last tile
            for (int e = tid; e < p.M * UROWS; e += kCT) stash[(e / UROWS) * TS + (e % UROWS)] = tile[(e / UROWS) * TS + (e % UROWS)];
            pend_rb = rb;
        } else {  // a lower-numbered share of rb: publish it for the reducer
            float* wsp = p.ws + (size_t)blockIdx.x * (8 * UROWS);
            for (int e = tid; e < p.M * UROWS; e += kCT) __stcg(wsp + e, tile[(e / UROWS) * TS + (e % UROWS)]);
            __threadfence();
            cbar();
            if (tid == 0) st_release_gpu(p.flags + blockIdx.x, 1u);
        }
        cbar();  // the tiles may be overwritten by the next flush
    };

    // ================================================================================================ main loop over x windows ======
    for (int wb = ubase; wb < u1; wb += p.xs) {
        const int wbeg = max(wb, u0), wend = min(wb + p.xs, u1);
        // ---- stage the activations of k steps (wb + slot) % ksteps, slot < xs, permuted to the fragment order ----------------------
        cbar();  // the previous window's activations are no longer read
        {
            const int nslot = wend - wb;
            for (int it = tid; it < nslot * MX * 4; it += kCT) {
                const int tt = it & 3, m = (it >> 2) % MX, slot = it / (4 * MX);
                const int ks = (wb + slot) % p.ksteps;
                uint32_t src[16];
                if (m < p.M && wb + slot >= u0) {
                    const size_t k0 = (size_t)ks * 128 + tt * 32;
                    const uint4* gp = reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + k0);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 v = __ldg(gp + i);
                        src[4 * i + 0] = v.x, src[4 * i + 1] = v.y, src[4 * i + 2] = v.z, src[4 * i + 3] = v.w;
                    }
                    if (p.norm_w) {  // (x * s) * w, rounded to bf16 like the stand-alone kernel's output
                        const float sc  = s_scale[m];
                        const uint4* wp = reinterpret_cast<const uint4*>(p.norm_w + k0);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
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
                    for (int i = 0; i < 16; i++) src[i] = 0u;
                }
                if (MODE == DM_FAST) {
                    // fp16 staging with a power-of-two scale per 128-k group (the quad's 4 x 32 values): the group's largest magnitude
                    // lands in [2^13, 2^14), so every bf16 input within 2^27 of it converts exactly; the elements that meet the codes
                    // carrying a factor 16 (even codes of every 8: the fp16 number built from the byte's high nibble) are divided by 16
                    const unsigned qm = 0xFu << (lane & ~3);
                    uint32_t amax = 0;
#pragma unroll
                    for (int i = 0; i < 16; i++) amax = max(amax, max(src[i] & 0x7fffu, (src[i] >> 16) & 0x7fffu));
                    amax = max(amax, __shfl_xor_sync(qm, amax, 1));
                    amax = max(amax, __shfl_xor_sync(qm, amax, 2));
                    const int e      = (int)(amax >> 7);                          // biased bf16 exponent of the largest magnitude
                    const int shift  = amax == 0 ? 0 : max(-100, min(100, 140 - e));  // 140 = 127 + 13
                    const float scl  = __uint_as_float((uint32_t)(127 + shift) << 23);
                    float sum = 0.f, sume = 0.f;  // Sx' (scaled) and Sxe (scaled, with the 1/16 factors as the MMA sees them)
                    uint32_t hsrc[16];            // fp16 pairs, natural order
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        // elements 2i (even index within its 8-block: code parity even -> factor 16 on the weight side) and 2i + 1
                        const float a = bf16lo(src[i]) * scl, b = bf16hi(src[i]) * scl;
                        const __half ha = __float2half_rn(a * 0.0625f), hb = __float2half_rn(b);
                        sum += a, sum += b;
                        sume += __half2float(ha), sume += __half2float(hb);
                        hsrc[i] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
                    }
                    sum += __shfl_xor_sync(qm, sum, 1), sum += __shfl_xor_sync(qm, sum, 2);      // fixed order: bit-reproducible
                    sume += __shfl_xor_sync(qm, sume, 1), sume += __shfl_xor_sync(qm, sume, 2);
                    if (tt == 0) sxs[slot * MX + m] = make_float4(fmaf(1024.0f, sume, (float)p.qbias * sum), sum, __uint_as_float((uint32_t)(127 - shift) << 23), 0.f);
#pragma unroll
                    for (int i = 0; i < 16; i++) src[i] = hsrc[i];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int e0 = xperm4(u * 8 + 2 * j), e1 = xperm4(u * 8 + 2 * j + 1);
                        const uint32_t lo = (src[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu;
                        const uint32_t hi = (src[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu;
                        o[j] = lo | (hi << 16);
                    }
                    xs[((slot * 4 + u) * MX + m) * 4 + tt] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        cbar();

        // ---- the stages of this window: 4 k steps each (never straddling a row block because K % 1024 == 0) ------------------------
        int uc = wbeg & ~(CHUNK - 1);
        int rb = uc / p.ksteps, ks0 = uc - rb * p.ksteps;  // the only division of the window; (rb, ks0) advance incrementally below
#pragma unroll 1
        for (; uc < wend; uc += CHUNK) {
            // ---- zero / step: a new tile every 8 k steps (and at the start of this CTA's share, which may sit mid-block) ----
            if ((ks0 & (GBLK - 1)) == 0 || uc == ubase) {
                if (uc != ubase) {  // the previous block is no longer read by this warp
                    __syncwarp();
                    if (lane == 0) mbar_arrive(gempty0 + 8 * gstage);
                    if (++gstage == kGStage) gstage = 0, gphase ^= 1;
                }
                mbar_wait(gfull0 + 8 * gstage, gphase);
            }
            const uint32_t ga = gring0 + gstage * GBYTES + (uint32_t)trow * 16 + (ks0 & (GBLK - 1)) * 2;  // zero of (row trow, group ks0)
            mbar_wait(full0 + 8 * stage, phase);
#pragma unroll
            for (int jp = 0; jp < UPW; jp += 2) {  // this warp group's k steps of the stage, two at a time
                const int j0 = UPW * wg + jp;
                uint32_t ra[2][4], rb8[2][4];
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    lds128(ring0 + stage * SBYTES + woff + (jp + jj) * UBYTES, ra[jj]);
                    lds128(ring0 + stage * SBYTES + woff + (jp + jj) * UBYTES + 512, rb8[jj]);
                }
                if (jp + 2 >= UPW) {  // last read of the stage: hand it back to the producer (one arrival per warp)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * stage);
                }
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const int u = uc + j0 + jj;
                    if (u < wbeg || u >= wend) continue;  // uniform over the warp group (ragged ends of the share)
#ifdef KF_DEBUG_KNOBS
                    if (p.dbg & 1) {
                        acc[0][0] += __uint_as_float(ra[jj][0] ^ rb8[jj][1]);
                        continue;
                    }
#endif
                    // zero / step of this k step's group for rows trow and trow + 8 (bf16 in the staged tile: [zero | step][128 rows][8 groups])
                    uint32_t za16, zb16, sa16, sb16;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(za16) : "r"(ga + (j0 + jj) * 2));
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(zb16) : "r"(ga + (j0 + jj) * 2 + 128));
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sa16) : "r"(ga + (j0 + jj) * 2 + GBYTES / 2));
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sb16) : "r"(ga + (j0 + jj) * 2 + GBYTES / 2 + 128));
                    const int slot = u - wb;
                    if (MODE == DM_FAST) {
                        // ---- fp16 codes 1024 + c / 1024 + 16 c straight into the tensor cores; affine map applied to the group sums ----
                        const float4 sv0 = sxs[slot * MX + ((2 * t) & (MX - 1))], sv1 = sxs[slot * MX + ((2 * t + 1) & (MX - 1))];
                        float accg[2][4];
                        accg[0][0] = accg[0][2] = -sv0.x, accg[0][1] = accg[0][3] = -sv1.x;  // -(1024 Sxe + qbias Sx)
#pragma unroll
                        for (int q = 0; q < 4; q++) accg[1][q] = 0.f;
#pragma unroll
                        for (int uu = 0; uu < 4; uu++) {
                            const uint4 xb = xs[((slot * 4 + uu) * MX) * 4 + xcol];
                            const uint32_t wa = ra[jj][3 - uu], wb8 = rb8[jj][3 - uu];  // the 128-bit words keep the first codes in the LAST register
                            const uint32_t wa_s = __byte_perm(wa, 0u, 0x4321), wb_s = __byte_perm(wb8, 0u, 0x4321);  // >> 8 as a byte permute
                            uint32_t a[4];
                            a[0] = and_or_imm<0x000F000Fu>(wa, magic), a[1] = and_or_imm<0x000F000Fu>(wb8, magic);   // codes {7, 3}
                            a[2] = and_or_imm<0x00F000F0u>(wa, magic), a[3] = and_or_imm<0x00F000F0u>(wb8, magic);   // codes {6, 2} x 16
                            mma_f16_16816(accg[0], a, xb.x, xb.y);
                            a[0] = and_or_imm<0x000F000Fu>(wa_s, magic), a[1] = and_or_imm<0x000F000Fu>(wb_s, magic);  // codes {5, 1}
                            a[2] = and_or_imm<0x00F000F0u>(wa_s, magic), a[3] = and_or_imm<0x00F000F0u>(wb_s, magic);  // codes {4, 0} x 16
                            mma_f16_16816(accg[1], a, xb.z, xb.w);
                        }
                        // y += 2^-shift * (step * sum((c - qbias) x') - zero * Sx')
                        const float fa = __uint_as_float(sa16 << 16), fb = __uint_as_float(sb16 << 16);
                        const float za = __uint_as_float(za16 << 16), zb = __uint_as_float(zb16 << 16);
                        acc[0][0] = fmaf(sv0.z, fmaf(fa, accg[0][0] + accg[1][0], -za * sv0.y), acc[0][0]);
                        acc[0][1] = fmaf(sv1.z, fmaf(fa, accg[0][1] + accg[1][1], -za * sv1.y), acc[0][1]);
                        acc[0][2] = fmaf(sv0.z, fmaf(fb, accg[0][2] + accg[1][2], -zb * sv0.y), acc[0][2]);
                        acc[0][3] = fmaf(sv1.z, fmaf(fb, accg[0][3] + accg[1][3], -zb * sv1.y), acc[0][3]);
                    } else {
                        // ---- the reference's per-weight dequant in bf16 (one or two roundings) ----
                        const uint32_t step2a = sa16 * 0x00010001u, step2b = sb16 * 0x00010001u;
                        uint32_t gza = za16 * 0x00010001u, gzb = zb16 * 0x00010001u;
                        if (MODE == DM_FMA) gza ^= 0x80008000u, gzb ^= 0x80008000u;  // -zero
#pragma unroll
                        for (int uu = 0; uu < 4; uu++) {
                            const uint4 xb = xs[((slot * 4 + uu) * MX) * 4 + xcol];
                            const uint32_t wa = ra[jj][3 - uu], wb8 = rb8[jj][3 - uu];
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                uint32_t a[4];
                                a[0] = deq_pair<MODE>(wa, 8 * h, step2a, gza, bias2, magic);
                                a[1] = deq_pair<MODE>(wb8, 8 * h, step2b, gzb, bias2, magic);
                                a[2] = deq_pair<MODE>(wa, 8 * h + 4, step2a, gza, bias2, magic);
                                a[3] = deq_pair<MODE>(wb8, 8 * h + 4, step2b, gzb, bias2, magic);
                                mma_bf16_16816(acc[h], a, h ? xb.z : xb.x, h ? xb.w : xb.y);  // two independent accumulation chains
                            }
                        }
                    }
                }
            }
            if (++stage == NS) stage = 0, phase ^= 1;
            // ---- end of a row block (always the end of a stage) or of this CTA's share: hand the sums over ----
            if (ks0 + CHUNK == p.ksteps || uc + CHUNK >= u1) {
                flush(rb, seg_k0, ks0 + min(CHUNK, u1 - uc));
                seg_k0 = 0;
            }
            ks0 += CHUNK;
            if (ks0 == p.ksteps) ks0 = 0, rb++;
        }
    }

    // ---- deferred stream-K reduction: this CTA holds the last share of row block pend_rb ---------------------------------------------
    if (pend_rb >= 0) {
        const int c_first = (pend_rb * p.ksteps) / p.upc, cnt = (int)blockIdx.x - c_first;  // lower shares: CTAs c_first .. blockIdx.x - 1
        if (tid < cnt) {
            const unsigned* f = p.flags + c_first + tid;
            unsigned spins = 0;
            while (ld_acquire_gpu(f) == 0u) {
                if (++spins > (1u << 26)) __trap();  // a lost contributor fails the launch instead of hanging the GPU
            }
        }
        cbar();
        for (int e = tid; e < p.M * UROWS; e += kCT) {
            float sum = 0.f;
            for (int c = 0; c < cnt; c++) sum += __ldcg(p.ws + (size_t)(c_first + c) * (8 * UROWS) + e);  // fixed order: bit-reproducible
            tile[(e / UROWS) * TS + (e % UROWS)] = sum + stash[(e / UROWS) * TS + (e % UROWS)];
        }
        cbar();
        if (tid < cnt) p.flags[c_first + tid] = 0u;  // self-reset for the next launch (which starts after this grid has completed)
        epilogue(pend_rb);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
struct MapKey {
    const void* base;
    int rows, cols;
    bool operator<(const MapKey& o) const { return base != o.base ? base < o.base : rows != o.rows ? rows < o.rows : cols < o.cols; }
};
struct TmaState {
    struct WMaps {
        CUtensorMap w, z, s;
    };
    std::map<MapKey, WMaps> maps;  // encoded once per weight
    float* ws       = nullptr;
    unsigned* flags = nullptr;
    int ws_ctas     = 0;
    uint64_t attr_devices[3 * 4 * 2] = {};  // per kernel instantiation: devices whose MaxDynamicSharedMemorySize has been set
};
TmaState* state_of(kf_ctx* ctx) {
    if (!ctx->gemv_tma) ctx->gemv_tma = new TmaState();
    return reinterpret_cast<TmaState*>(ctx->gemv_tma);
}

template <int MODE, int MX, int NWG>
int launch(kf_ctx* ctx, TmaState* st, const TParams& p, const TMaps& tm, int grid, size_t smem) {
    auto kern = kf_gemv_tma_kernel<MODE, MX, NWG>;
    constexpr int slot = (MODE * 4 + (MX == 1 ? 0 : MX == 2 ? 1 : MX == 4 ? 2 : 3)) * 2 + (NWG - 1);
    if (!(st->attr_devices[slot] >> (ctx->device & 63) & 1)) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        st->attr_devices[slot] |= 1ull << (ctx->device & 63);
    }
    KF_CUDA(ctx, kf_launch_pdl(ctx, kern, dim3(grid), dim3(256 * NWG + 32), smem, tm, p));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
template <int MODE, int NWG>
int launch_mx(kf_ctx* ctx, TmaState* st, const TParams& p, const TMaps& tm, int grid, size_t smem, int mx) {
    switch (mx) {
        case 1: return launch<MODE, 1, NWG>(ctx, st, p, tm, grid, smem);
        case 2: return launch<MODE, 2, NWG>(ctx, st, p, tm, grid, smem);
        case 4: return launch<MODE, 4, NWG>(ctx, st, p, tm, grid, smem);
        default: return launch<MODE, 8, NWG>(ctx, st, p, tm, grid, smem);
    }
}
template <int MODE>
int launch_wg(kf_ctx* ctx, TmaState* st, const TParams& p, const TMaps& tm, int grid, size_t smem, int mx, int nwg) {
    return nwg == 2 ? launch_mx<MODE, 2>(ctx, st, p, tm, grid, smem, mx) : launch_mx<MODE, 1>(ctx, st, p, tm, grid, smem, mx);
}
}  // namespace

void kf_gemv_tma_destroy(kf_ctx* ctx) {
    TmaState* st = reinterpret_cast<TmaState*>(ctx->gemv_tma);
    if (!st) return;
    if (st->ws) cudaFree(st->ws);
    if (st->flags) cudaFree(st->flags);
    delete st;
    ctx->gemv_tma = nullptr;
}

int kf_make_tensor_map_2d(kf_ctx* ctx, CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner, uint64_t outer,
                          uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
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
    return KF_OK;
}

// Returns KF_OK when the launch was issued, 1 when this kernel does not cover the request (the caller falls back to gemv.cu), < 0 on error.
int kf_gemv_tma(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                const void* norm_w, float norm_eps) {
    if (!ctx->gemv_tma_on || n < 1 || n > 3 || M < 1 || M > 8) return 1;
    const int K = w[0].cols;
    if (w[0].type != KF_T_Q4 || K % (128 * GBLK) != 0 || w[0].group != 128) return 1;
    for (int i = 0; i < n; i++) {
        if (w[i].type != KF_T_Q4 || w[i].cols != K || w[i].group != 128 || w[i].qbias != w[0].qbias || !kf_has_gama(w[i])) return 1;
        if (w[i].rows % 16 != 0 || w[i].rows < 16 || ((uintptr_t)w[i].data_dev & 15) || !y[i]) return 1;
        if (((uintptr_t)kf_gama_zero(w[i]) & 15) || ((uintptr_t)kf_gama_step(w[i]) & 15)) return 1;  // TMA tiles of 8 groups = 16 bytes
    }
    if (epilogue == EPI_SWIGLU && (n != 2 || w[0].rows != w[1].rows)) return 1;
    if (epilogue == EPI_RESIDUAL && (n != 1 || !residual)) return 1;
    if (((uintptr_t)x & 15) || (norm_w && ((uintptr_t)norm_w & 15))) return 1;
    TmaState* st = state_of(ctx);

    TParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = n, p.x = (const uint16_t*)x, p.residual = (const uint16_t*)residual, p.norm_w = (const uint16_t*)norm_w, p.norm_eps = norm_eps;
    p.M = M, p.K = K, p.ksteps = K / 128, p.qbias = w[0].qbias, p.epilogue = epilogue;
    p.dbg = (ctx->debug_skip >> 4) & 0xff;
    TMaps tm;
    int rb = 0;
    for (int i = 0; i < n; i++) {
        p.seg[i].zero = kf_gama_zero(w[i]), p.seg[i].step = kf_gama_step(w[i]), p.seg[i].y = (uint16_t*)y[i], p.seg[i].rows = w[i].rows;
        p.seg[i].rb0 = rb;
        rb += (w[i].rows + UROWS - 1) / UROWS;
        const MapKey key{w[i].data_dev, w[i].rows, K};
        auto it = st->maps.find(key);
        if (it == st->maps.end()) {
            TmaState::WMaps m;
            int rc = kf_make_tensor_map_2d(ctx, &m.w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, w[i].data_dev, (uint64_t)K / 2, (uint64_t)w[i].rows, 64, 64,
                                           CU_TENSOR_MAP_SWIZZLE_NONE);
            if (!rc)
                rc = kf_make_tensor_map_2d(ctx, &m.z, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, p.seg[i].zero, (uint64_t)K / 128, (uint64_t)w[i].rows, GBLK, 64,
                                           CU_TENSOR_MAP_SWIZZLE_NONE);
            if (!rc)
                rc = kf_make_tensor_map_2d(ctx, &m.s, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, p.seg[i].step, (uint64_t)K / 128, (uint64_t)w[i].rows, GBLK, 64,
                                           CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
            if (st->maps.size() > 65536) st->maps.clear();
            it = st->maps.emplace(key, m).first;
        }
        tm.w[i] = it->second.w, tm.z[i] = it->second.z, tm.s[i] = it->second.s;
    }
    for (int i = n; i < 3; i++) tm.w[i] = tm.w[0], tm.z[i] = tm.z[0], tm.s[i] = tm.s[0];
    if (epilogue == EPI_SWIGLU) rb = (w[0].rows + 63) / 64;
    p.total_rb = rb;
    p.units    = rb * p.ksteps;

    // launch geometry.  gemv_tma_warps = 8 (default): one CTA of 8 consumer warps + the producer per SM and HALF an SM's shared memory, so
    // that the next kernel of the stream becomes resident beside it; 16: two warp groups; gemv_tma_occ = 2: two CTAs per SM of the same launch
    const int mx   = M == 1 ? 1 : M == 2 ? 2 : M <= 4 ? 4 : 8;
    const int mode = !ctx->gemv_exact ? DM_FAST : ctx->deq_fma ? DM_FMA : DM_TWO;
    const int nwg  = ctx->gemv_tma_warps >= 16 ? 2 : 1;
    const int occ  = (ctx->gemv_tma_occ >= 2 && nwg == 1) ? 2 : 1;
    int grid       = std::min(p.units, ctx->sm_count * occ);
    p.upc          = (p.units + grid - 1) / grid;
    grid           = (p.units + p.upc - 1) / p.upc;
    // shared memory: weight ring | zero / step ring | x window | group sums | tiles | stash | barriers
    const size_t budget = (size_t)std::max(72, std::min(ctx->gemv_tma_smem_kb, 220)) * 1024;
    int xs_cap = (int)std::min<size_t>(64, (32 * 1024) / ((size_t)mx * 256));
    int xs     = std::min(xs_cap, ((p.upc + 3 + 3) & ~3));  // a CTA's share starts up to 3 steps after a 4-aligned slot
    xs         = std::max(4, xs & ~3);
    const size_t xbytes = (size_t)xs * mx * 256, sxbytes = mode == DM_FAST ? (size_t)xs * mx * 16 : 0;
    const size_t tileb  = ((size_t)mx * TS * 4 + 15) & ~(size_t)15;
    const size_t barb   = 2 * 8 * kMaxStage + 2 * 8 * kGStage;
    const size_t fixed  = (size_t)kGStage * GBYTES + xbytes + sxbytes + (nwg + 1) * tileb + barb + 64;
    if (fixed + 2 * SBYTES > budget) return 1;
    int ns = (int)std::min<size_t>(kMaxStage, (budget - fixed) / SBYTES);
    ns     = std::min(ns, std::max(2, (p.upc + 2 * CHUNK - 1) / CHUNK));
    p.nstage = ns, p.xs = xs;
    p.off_g     = ns * SBYTES;
    p.off_x     = p.off_g + kGStage * GBYTES;
    p.off_sx    = p.off_x + (int)xbytes;
    p.off_tile  = p.off_sx + (int)sxbytes;
    p.off_stash = p.off_tile + (int)(nwg * tileb);
    p.off_bar   = p.off_stash + (int)tileb;
    const size_t smem = (size_t)p.off_bar + barb;

    if (grid > st->ws_ctas) {  // stream-K workspace: one partial tile + one flag per CTA (sized once: grid <= SMs x occupancy)
        KF_REQUIRE(ctx, !ctx->capturing, "stream-K workspace must be sized before graph capture (run one eager step first)");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (st->ws) cudaFree(st->ws);
        if (st->flags) cudaFree(st->flags);
        st->ws = nullptr, st->flags = nullptr;
        const int want = std::max(grid, ctx->sm_count * 2);
        KF_CUDA(ctx, cudaMalloc(&st->ws, (size_t)want * 8 * UROWS * sizeof(float)));
        KF_CUDA(ctx, cudaMalloc(&st->flags, (size_t)want * sizeof(unsigned)));
        KF_CUDA(ctx, cudaMemsetAsync(st->flags, 0, (size_t)want * sizeof(unsigned), ctx->stream));
        st->ws_ctas = want;
        ctx->scratch_gen++;
    }
    p.ws = st->ws, p.flags = st->flags;
    if (mode == DM_FMA) return launch_wg<DM_FMA>(ctx, st, p, tm, grid, smem, mx, nwg);
    if (mode == DM_TWO) return launch_wg<DM_TWO>(ctx, st, p, tm, grid, smem, mx, nwg);
    return launch_wg<DM_FAST>(ctx, st, p, tm, grid, smem, mx, nwg);
}
