// ref_kernels_q.cu -- TEST INFRASTRUCTURE ONLY (see koifish_oracle.h): the reference's quantizer translation unit
// (src/Device/CUDA/kernel/quantizer.cu) compiled where it lies, to run its NormalFloat4 dequant kernel CU_Q42X_NF4 (:612-654) on the B200
// next to kf_dequant.  Only the kernel is used; the host functions of that file (GTensor::GetDataX ...) reference the rest of the
// framework, so their undefined symbols are bound to address 0 at link time (oracle/Makefile, --defsym) and never called.
#include "Device/CUDA/kernel/quantizer.cu"

// grid = rows, block = a divisor of cols / 2 (Q_nThreadOfBlock's role, cuda_def.hpp): every thread takes cols / block elements
extern "C" int refq_nf4_dequant(const void* gama_dev, const void* data_dev, void* out_bf16_dev, int rows, int cols) {
    int threads = 256;
    while (threads > 1 && (cols % (2 * threads)) != 0) threads /= 2;
    CU_Q42X_NF4<bf16><<<rows, threads>>>((floatGama*)gama_dev, (hBITARR)data_dev, (bf16*)out_bf16_dev, rows, cols, 0, 42);
    return (int)cudaDeviceSynchronize();
}

// CU_Q42X_awq (quantizer.cu:132-156) on a zero-filled TASKA_quant carrying only what the kernel reads (nOut = rows M, nIn = columns N of the
// stored [M][N] matrix); grid = M rows, block = threads with N / threads a multiple of 8
extern "C" int refq_awq_dequant(const void* qzeros_dev, const void* scales_dev, const void* qweight_dev, void* out_bf16_dev, int M, int N) {
    alignas(16) unsigned char raw[sizeof(TASKA_quant<bf16>)];
    memset(raw, 0, sizeof(raw));
    TASKA_quant<bf16>& t = *reinterpret_cast<TASKA_quant<bf16>*>(raw);
    t.nOut = M, t.nIn = N;
    int threads = 256;
    while (threads > 1 && (N % (8 * threads)) != 0) threads /= 2;
    CU_Q42X_awq<bf16><<<M, threads>>>(t, (const Q4_8*)qzeros_dev, (const half*)scales_dev, (const Q4_8*)qweight_dev, (bf16*)out_bf16_dev, 0);
    return (int)cudaDeviceSynchronize();
}
