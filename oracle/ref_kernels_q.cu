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
