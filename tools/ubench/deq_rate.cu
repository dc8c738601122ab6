// deq_rate.cu -- microbenchmark: cycles per "unit" (2 rows x 32 weights per thread: the work of one 128-row x 128-k tile for one warp)
// of the dequant + mma.sync instruction stream, from registers only (no memory), for 1..6 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o deq_rate deq_rate.cu && ./deq_rate
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ __nv_bfloat162 as2(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t asu(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// MODE 0: exact fma (sub + fma), 1: factor (no fp math), 2: factor without the MMA (xor accumulate), 3: exact without the MMA
template <int MODE>
__device__ __forceinline__ uint32_t deq(uint32_t reg, int shift, uint32_t step2, uint32_t gz, uint32_t bias2, uint32_t mask, uint32_t magic) {
    uint32_t v = and_or(reg >> shift, mask, magic);
    if (MODE == 1 || MODE == 2) return v;
    __nv_bfloat162 k = __hsub2_rn(as2(v), as2(bias2));
    return asu(__hfma2(k, as2(step2), as2(gz)));
}

template <int MODE>
__global__ void kern(uint32_t* out, long long* cyc, int iters, uint32_t seed, uint32_t mask, uint32_t magic) {
    uint32_t ra[4], rb[4], xb[4][4];
    for (int i = 0; i < 4; i++) ra[i] = seed * (threadIdx.x + 1 + i), rb[i] = seed * (threadIdx.x + 77 + i);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) xb[i][j] = 0x3f803f80u + i + j;
    float acc[2][4] = {};
    uint32_t step2 = 0x3c003c00u, gz = 0xbd00bd00u, bias2 = 0x43004300u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int uu = 0; uu < 4; uu++) {
            const uint32_t wa = ra[3 - uu], wb = rb[3 - uu];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t a[4];
                a[0] = deq<MODE>(wa, 8 * h, step2, gz, bias2, mask, magic);
                a[1] = deq<MODE>(wb, 8 * h, step2, gz, bias2, mask, magic);
                a[2] = deq<MODE>(wa, 8 * h + 4, step2, gz, bias2, mask, magic);
                a[3] = deq<MODE>(wb, 8 * h + 4, step2, gz, bias2, mask, magic);
                if (MODE <= 1)
                    mma(acc[h], a, h ? xb[uu][2] : xb[uu][0], h ? xb[uu][3] : xb[uu][1]);
                else {
                    acc[h][0] = __uint_as_float(__float_as_uint(acc[h][0]) ^ a[0] ^ a[1]);
                    acc[h][1] = __uint_as_float(__float_as_uint(acc[h][1]) ^ a[2] ^ a[3]);
                }
            }
        }
        // fresh "weights" for the next unit (cheap, keeps the compiler from hoisting)
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

template <int MODE>
void run(const char* name) {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16, 24, 32}) {
        const int iters = 2000;
        kern<MODE><<<148, warps * 32>>>(out, cyc, 10, 12345u, 0x000F000Fu, 0x43004300u);
        kern<MODE><<<148, warps * 32>>>(out, cyc, iters, 12345u, 0x000F000Fu, 0x43004300u);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; i++) c += (double)h[i];
        c /= 148.0 * iters;
        printf("%-28s warps/SM %2d : %7.1f cycles per unit per warp, %6.1f cycles per tile (8 warp-units) per SM\n", name, warps, c, c * 8.0 / warps);
    }
    cudaFree(out), cudaFree(cyc);
}
int main() {
    run<0>("exact (sub+fma) + mma");
    run<1>("factor + mma");
    run<3>("exact, no mma");
    run<2>("factor, no mma");
    return 0;
}
