// deq_rate2.cu -- which form of the unpack is cheapest on the ALU pipe?  Same harness as deq_rate.cu (registers only), 16 / 24 warps per SM.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (a & b) | c with b, c in registers (3 register reads)
__device__ __forceinline__ uint32_t lop_rrr(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// (a & IMM) | c : mask as an immediate (2 register reads)
template <uint32_t IMM>
__device__ __forceinline__ uint32_t lop_rir(uint32_t a, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c));
    return d;
}
__device__ __forceinline__ __nv_bfloat162 as2(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t asu(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// V: 0 = factor, 3-register LOP3 (today)   1 = factor, immediate mask   2 = factor, fp16 trick (2 shifts per 8 codes, immediate masks)
//    3 = exact (sub + fma), immediate mask  4 = exact, 3-register LOP3 (today)
template <int V>
__global__ void kern(uint32_t* out, long long* cyc, int iters, uint32_t seed, uint32_t mask, uint32_t magic) {
    uint32_t ra[4], rb[4], xb[4][4];
    for (int i = 0; i < 4; i++) ra[i] = seed * (threadIdx.x + 1 + i), rb[i] = seed * (threadIdx.x + 77 + i);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) xb[i][j] = 0x3f803f80u + i + j;
    float acc[2][4] = {};
    const uint32_t step2 = 0x3c003c00u, gz = 0xbd00bd00u, bias2 = 0x43004300u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int uu = 0; uu < 4; uu++) {
            const uint32_t wa = ra[3 - uu], wb = rb[3 - uu];
            if (V == 2) {
                // 8 codes of wa / wb: pairs (0,4) (1,5) from the register as is, (2,6) (3,7) from (reg >> 8); codes 1,5 / 3,7 carry a factor 16
                const uint32_t wa8 = wa >> 8, wb8 = wb >> 8;
                uint32_t a[4];
                a[0] = lop_rir<0x000F000Fu>(wa, magic), a[1] = lop_rir<0x000F000Fu>(wb, magic);
                a[2] = lop_rir<0x00F000F0u>(wa, magic), a[3] = lop_rir<0x00F000F0u>(wb, magic);
                mma_f16(acc[0], a, xb[uu][0], xb[uu][1]);
                a[0] = lop_rir<0x000F000Fu>(wa8, magic), a[1] = lop_rir<0x000F000Fu>(wb8, magic);
                a[2] = lop_rir<0x00F000F0u>(wa8, magic), a[3] = lop_rir<0x00F000F0u>(wb8, magic);
                mma_f16(acc[1], a, xb[uu][2], xb[uu][3]);
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t a[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t r = (q & 1) ? wb : wa;
                        const int sh = 8 * h + 4 * (q >> 1);
                        uint32_t v = (V == 0 || V == 4) ? lop_rrr(r >> sh, mask, magic) : lop_rir<0x000F000Fu>(r >> sh, magic);
                        if (V >= 3) {
                            __nv_bfloat162 k = __hsub2_rn(as2(v), as2(bias2));
                            v = asu(__hfma2(k, as2(step2), as2(gz)));
                        }
                        a[q] = v;
                    }
                    mma_bf16(acc[h], a, h ? xb[uu][2] : xb[uu][0], h ? xb[uu][3] : xb[uu][1]);
                }
            }
        }
        ra[it & 3] += __float_as_uint(acc[0][0]) | 1u;
        rb[(it + 1) & 3] ^= __float_as_uint(acc[1][1]);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int h = 0; h < 2; h++)
        for (int j = 0; j < 4; j++) s ^= __float_as_uint(acc[h][j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name) {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int warps : {8, 16, 24}) {
        const int iters = 2000;
        kern<V><<<148, warps * 32>>>(out, cyc, 10, 12345u, 0x000F000Fu, V == 2 ? 0x64006400u : 0x43004300u);
        kern<V><<<148, warps * 32>>>(out, cyc, iters, 12345u, 0x000F000Fu, V == 2 ? 0x64006400u : 0x43004300u);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; i++) c += (double)h[i];
        c /= 148.0 * iters;
        printf("%-44s warps/SM %2d : %6.1f cycles per tile (8 warp-units) per SM\n", name, warps, c * 8.0 / warps);
    }
    cudaFree(out), cudaFree(cyc);
}
int main() {
    run<0>("factor, LOP3 r,r,r (round 1 form)");
    run<1>("factor, LOP3 r,imm,r");
    run<2>("factor, fp16 trick: 2 SHF + 8 LOP3 per 8 pairs");
    run<4>("exact sub+fma, LOP3 r,r,r");
    run<3>("exact sub+fma, LOP3 r,imm,r");
    return 0;
}
