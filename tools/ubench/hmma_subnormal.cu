// Does mma.sync.m16n8k16.f32.f16.f16.f32 on sm_100a treat fp16 SUBNORMAL A operands exactly (no flush to zero)?
// A[i][k] = m * 2^-24 (exponent field 0, mantissa m in 0..1023: what `packed_word & mask` yields for a code field inside the low 10 bits),
// B = fp16 activations up to 2^14.  Compared with the double-precision sum.   nvcc -arch=sm_100a -o hmma_subnormal hmma_subnormal.cu
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k(const uint16_t* A, const uint16_t* B, float* D) {  // A [16][16] row-major, B [16 k][8 n] as B[k][n], D [16][8]
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    auto pk = [](uint16_t lo, uint16_t hi) { return (uint32_t)lo | ((uint32_t)hi << 16); };
    uint32_t a[4], b[2];
    a[0] = pk(A[g * 16 + 2 * t], A[g * 16 + 2 * t + 1]);
    a[1] = pk(A[(g + 8) * 16 + 2 * t], A[(g + 8) * 16 + 2 * t + 1]);
    a[2] = pk(A[g * 16 + 2 * t + 8], A[g * 16 + 2 * t + 9]);
    a[3] = pk(A[(g + 8) * 16 + 2 * t + 8], A[(g + 8) * 16 + 2 * t + 9]);
    b[0] = pk(B[(2 * t) * 8 + g], B[(2 * t + 1) * 8 + g]);
    b[1] = pk(B[(2 * t + 8) * 8 + g], B[(2 * t + 9) * 8 + g]);
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    D[g * 8 + 2 * t] = d[0], D[g * 8 + 2 * t + 1] = d[1], D[(g + 8) * 8 + 2 * t] = d[2], D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

int main() {
    uint16_t hA[256], hB[128];
    float hD[128];
    uint16_t *dA, *dB;
    float* dD;
    cudaMalloc(&dA, sizeof hA), cudaMalloc(&dB, sizeof hB), cudaMalloc(&dD, sizeof hD);
    double worst = 0;
    srand(1);
    for (int trial = 0; trial < 200; trial++) {
        const int kind = trial % 4;  // 0: 4-bit code at bit 0, 1: 4-bit code at bit 4, 2: 2-bit fields at 0..8, 3: any 10-bit mantissa
        for (int i = 0; i < 256; i++) {
            int c = rand();
            hA[i] = kind == 0 ? (c & 15) : kind == 1 ? ((c & 15) << 4) : kind == 2 ? ((c & 3) << (2 * ((c >> 8) % 5))) : (c & 1023);
        }
        for (int i = 0; i < 128; i++) {
            float x = ((rand() % 20001) - 10000) / 10000.0f * ldexpf(1.0f, (rand() % 28) - 13);  // magnitudes 2^-13 .. 2^14
            hB[i]   = __half_as_ushort(__float2half_rn(x));
        }
        cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice), cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(dA, dB, dD);
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 8; j++) {
                double ref = 0, mag = 0;
                for (int kk = 0; kk < 16; kk++) {
                    double p = (double)hA[i * 16 + kk] * ldexp(1.0, -24) * (double)__half2float(__ushort_as_half(hB[kk * 8 + j]));
                    ref += p, mag += fabs(p);
                }
                double err = fabs((double)hD[i * 8 + j] - ref) / (mag > 0 ? mag : 1);
                if (err > worst) worst = err;
            }
    }
    printf("fp16 subnormal A operands through mma.sync.m16n8k16 f16 -> f32: worst |D - exact| / sum|products| = %.3e  (%s)\n", worst,
           worst < 1e-6 ? "subnormals are exact: not flushed" : "NOT exact");
    return cudaGetLastError() != cudaSuccess;
}
