// hmma_rate.cu -- microbenchmark: issue rate of the legacy mma.sync.m16n8k16 bf16 path on sm_100a, per SM, for 1/2/4/8
// independent accumulator chains per warp and 4..16 warps per SM.  Explains the ceiling of the first GEMV versions
// (profiles/r01_gemv_sweep_v2.jsonl: ~13 cycles per HMMA per SM regardless of bit width).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void hmma_kernel(float* out, int iters, unsigned a0) {
    float acc[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
        for (int j = 0; j < 4; j++) acc[c][j] = 0.f;
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b0 = a0 ^ 0x3c003c00u, b1 = a0 ^ 0x3c013c01u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
        for (int j = 0; j < 4; j++) s += acc[c][j];
    if (s == 123.456f) out[0] = s;
}

template <int CHAINS>
void run(int warps) {
    int sms = 148;
    float* out;
    cudaMalloc(&out, 4);
    int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    hmma_kernel<CHAINS><<<sms, warps * 32>>>(out, 64, 0x3f803f80u);
    cudaEventRecord(e0);
    hmma_kernel<CHAINS><<<sms, warps * 32>>>(out, iters, 0x3f803f80u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double hmma_per_sm = (double)iters * CHAINS * warps;
    double cycles      = ms * 1e-3 * 1.965e9;
    printf("chains %d warps/SM %2d : %.2f cycles per HMMA per SM  (%.1f dense TFLOP/s chip)\n", CHAINS, warps, cycles / hmma_per_sm,
           hmma_per_sm * sms * 4096.0 * 2 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<1>(w);
        run<2>(w);
        run<4>(w);
        run<8>(w);
    }
    return 0;
}
