// ref_kernels.cu -- TEST INFRASTRUCTURE ONLY: the reference's OWN CUDA kernels for the hot path, compiled for sm_100a from the
// sources where they lie under /root/reference (never copied), and exposed through a flat C interface so that `-m gpu` parity tests
// can run them on the B200 next to koifish_b200's kernels.  Built by oracle/Makefile (target `refgpu`) into
// oracle/_ref/libkoifish_refgpu.so; the product (koifish_b200/) never loads it.
//
// What is pinned (reference file:line):
//   refk_q128tox           CU_Q128toX_<bf16,32|64|128>     src/Device/CUDA/T.cu:245-294     (GetDataX for Q4 / T_SIGN / T_BINARY)
//   refk_xtoq128           CU_XtoQ128_<bf16,32|64>          src/Device/CUDA/T.cu:105-174     (SetDataX quantise+pack, 4- / 2-bit)
//                          CU_XtoYYang_<bf16>               src/Device/CUDA/T.cu:176-242     (1-bit)
//   refk_rmsnorm           rms_norm_kernel<256> (CU_rms_infer)  src/Device/CUDA/kernel/layernorm.cuh:801-859
//   refk_rmsnorm_multihead CU_rmsnorm_multihead             src/Device/CUDA/kernel/layernorm.cuh:750-798
//   refk_rope2             CU_rope2_v0                      src/Device/CUDA/kernel/operator.cuh:735-772  (stochastic rounding, seed 42)
//   refk_attention         attention_qk_kernel + CU_softmax_multihead + attention_v_kernel   operator.cuh:573-632, 252-277, 650-668
//                          launched as SelfAttention::cuInfer does (src/Device/CUDA/QKV.cu:667-672)
//   refk_f8_decode/encode  CU_F82Float / CU_Float2F8<bf16>  src/Device/CUDA/kernel/operator.cuh:519-543
// The launch geometry is the reference's (TASKA_quant BLOCK_at_GROUP: 128 threads per block, one thread per group,
// src/Tensor/GeQuant.cpp:1297-1350).  TASKA_quant has no default constructor and its real constructors live in GeQuant.cpp (which only
// links with the whole framework), so the struct is zero-filled and the fields the kernels read are set by hand.
//
// The whole translation unit T.cu is included (the kernel templates are defined there, not in a header); the handful of host symbols
// its non-kernel code references are defined at the bottom as inert stand-ins -- none of that host code is ever called.
#include "Device/CUDA/T.cu"

#include <new>

namespace {
template <class K, class... A>
int launch(K kernel, dim3 grid, dim3 block, A... args) {
    kernel<<<grid, block>>>(args...);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : -(int)e;
}
struct TaskBox {
    alignas(16) unsigned char raw[sizeof(TASKA_quant<floatX>)];
    TASKA_quant<floatX>& t() { return *reinterpret_cast<TASKA_quant<floatX>*>(raw); }
    TaskBox(int nG, int lG, int qMin, int qMax, int qBias, int isSym, int yyang, void* zero, void* step) {
        memset(raw, 0, sizeof(raw));
        TASKA_quant<floatX>& q = t();
        q.nG = nG, q.lG = lG, q.qMin = qMin, q.qMax = qMax, q.qBias = qBias, q.isSym = isSym != 0;
        q.yyang = (QUANT_YYANG_)yyang, q.rc_normal = 0, q.isAccumErr = false, q.seed = 42;
        q.zero = (floatGama*)zero, q.step = (floatGama*)step;
        q.distill.lenda = -1.f, q.distill.lendaW = nullptr;
        q.tpb = q.block3 = 128, q.nBlock = q.grid3 = (nG + 127) / 128;  // BLOCK_at_GROUP, GeQuant.cpp:1309-1312
    }
};
}  // namespace

extern "C" {
// yyang: 0 I_OFF, 1 I_01 (1-bit), 3 I_TERNARY (2-bit)   (QUANT_YYANG_, src/CLI_params.hpp:502-507)
int refk_q128tox(int bits, int nG, int lG, int qbias, const void* packed_dev, const void* zero_dev, const void* step_dev, void* out_bf16_dev) {
    TaskBox b(nG, lG, 0, 0, qbias, 0, 0, (void*)zero_dev, (void*)step_dev);
    const dim3 grid(b.t().grid3), block(b.t().block3);
    if (bits == 4) return launch(CU_Q128toX_<floatX, 32>, grid, block, b.t(), (const BIT_128*)packed_dev, (floatX*)out_bf16_dev, 0);
    if (bits == 2) return launch(CU_Q128toX_<floatX, 64>, grid, block, b.t(), (const BIT_128*)packed_dev, (floatX*)out_bf16_dev, 0);
    if (bits == 1) return launch(CU_Q128toX_<floatX, 128>, grid, block, b.t(), (const BIT_128*)packed_dev, (floatX*)out_bf16_dev, 0);
    return -1;
}
int refk_xtoq128(int bits, int nG, int lG, int qMin, int qMax, int qBias, int isSym, int yyang, const void* in_bf16_dev, void* packed_dev,
                 void* zero_dev, void* step_dev) {
    TaskBox b(nG, lG, qMin, qMax, qBias, isSym, yyang, zero_dev, step_dev);
    const dim3 grid(b.t().grid3), block(b.t().block3);
    if (bits == 4) return launch(CU_XtoQ128_<floatX, 32>, grid, block, b.t(), (BIT_128*)packed_dev, (const floatX*)in_bf16_dev, 0);
    if (bits == 2) return launch(CU_XtoQ128_<floatX, 64>, grid, block, b.t(), (BIT_128*)packed_dev, (const floatX*)in_bf16_dev, 0);
    if (bits == 1) return launch(CU_XtoYYang_<floatX>, grid, block, b.t(), (BIT_128*)packed_dev, (const floatX*)in_bf16_dev, 0);
    return -1;
}
// CU_rms_infer: one block of 256 threads per row, eps = the kernel's default 1e-6 (layernorm.cuh:849-859)
int refk_rmsnorm(void* out_dev, const void* x_dev, const void* w_dev, int rows, int dim) {
    for (int r = 0; r < rows; r++) {
        int rc = launch(rms_norm_kernel<CU_T4B_SMALL, floatX>, dim3(1), dim3(CU_T4B_SMALL), (floatX*)out_dev + (size_t)r * dim,
                        (const floatX*)x_dev + (size_t)r * dim, (const floatX*)w_dev, (size_t)dim, 1.0f / dim, 1e-6f);
        if (rc) return rc;
    }
    return 0;
}
int refk_rmsnorm_multihead(void* vecs_dev, const void* w_dev, int n_head, int head_dim, int threads) {
    return launch(CU_rmsnorm_multihead, dim3(n_head), dim3(threads), (bf16*)vecs_dev, (const bf16*)w_dev, n_head, head_dim, 1e-6f);
}
// ROPE::cuInfer, fuse_normal == 0 branch (rope.cu:666): CU_rope2_v0<<<(B,T,n_head), head_dim/2>>>(q, k, pos, ..., theta, 42)
int refk_rope2(void* q_dev, void* k_dev, int pos, int n_head, int n_kv, int head_dim, float theta) {
    return launch(CU_rope2_v0<floatX>, dim3(1, 1, n_head), dim3(head_dim / 2), (floatX*)q_dev, (floatX*)k_dev, pos, n_head, n_kv, head_dim, theta,
                  42, 0);
}
// score_bf16 = 1: the neuron path's bf16 score buffer (qk_v is tpWeight, src/Manifold/TGraph.cpp:123-124); 0: the pipe path's fp32 buffer
int refk_attention(void* out_dev, void* att_scratch_dev, const void* q_dev, const void* kcache_dev, const void* vcache_dev, int pos, int seq_len,
                   int n_head, int n_kv, int head_dim, int score_bf16) {
    const int thr = pos + 1 < 1024 ? pos + 1 : 1024;  // QKV.cu:667
    int rc;
    if (score_bf16) {
        bf16* att = (bf16*)att_scratch_dev;
        rc = launch(attention_qk_kernel<bf16>, dim3(n_head), dim3(thr), att, (bf16*)q_dev, (bf16*)kcache_dev, pos, seq_len, n_head, n_kv, head_dim);
        if (!rc) rc = launch(CU_softmax_multihead<bf16>, dim3(n_head), dim3(1), att, pos, seq_len);
        if (!rc)
            rc = launch(attention_v_kernel<bf16>, dim3(n_head), dim3(head_dim), (bf16*)out_dev, (const bf16*)att, (const bf16*)vcache_dev, pos, seq_len,
                        n_head, n_kv, head_dim);
    } else {
        float* att = (float*)att_scratch_dev;
        rc = launch(attention_qk_kernel<float>, dim3(n_head), dim3(thr), att, (bf16*)q_dev, (bf16*)kcache_dev, pos, seq_len, n_head, n_kv, head_dim);
        if (!rc) rc = launch(CU_softmax_multihead<float>, dim3(n_head), dim3(1), att, pos, seq_len);
        if (!rc)
            rc = launch(attention_v_kernel<float>, dim3(n_head), dim3(head_dim), (bf16*)out_dev, (const float*)att, (const bf16*)vcache_dev, pos,
                        seq_len, n_head, n_kv, head_dim);
    }
    return rc;
}
int refk_f8_decode(const void* f8_dev, void* out_bf16_dev, size_t n) {
    return launch(CU_F82Float<floatX>, dim3((unsigned)((n + CU_T4B_MIDDLE - 1) / CU_T4B_MIDDLE)), dim3(CU_T4B_MIDDLE), (const f8e5*)f8_dev,
                  (floatX*)out_bf16_dev, n, 0, 0);
}
int refk_f8_encode(const void* in_bf16_dev, void* f8_dev, size_t n) {
    return launch(CU_Float2F8<floatX>, dim3((unsigned)((n + CU_T4B_MIDDLE - 1) / CU_T4B_MIDDLE)), dim3(CU_T4B_MIDDLE), (const floatX*)in_bf16_dev,
                  (f8e5*)f8_dev, n, 0, 0);
}
int refk_sizeof_taska(void) { return (int)sizeof(TASKA_quant<floatX>); }
}

// ---- inert stand-ins for the host symbols referenced by T.cu's non-kernel code (GTensor::SetDataX, huTensor::Quant4A, the layernorm
//      host wrappers): that code is compiled because the TU is included whole, but nothing here ever calls it ----
int g_dump_level = 0, g_dump_each = 0, g_dump_sigfigs = 6;
cudaDeviceProp deviceProp;
void* GTensor::buff     = nullptr;
size_t GTensor::buff_len = 0;
bool D2H(const void*, void*, size_t, int) { return false; }
void _LOG(DUMP_LEVEL, const char*, ...) {}
const char* cNameOf(typNUMBER) { return "?"; }
bool Fish::isAtPhase(LIFE_PHASE) const { return false; }
cudaStream_t main_stream = nullptr;
// the real constructors live in src/Tensor/GeQuant.cpp:1297-1372 (they walk GTensor / GeQuant objects); never called here
template <typename Typ>
TASKA_quant<Typ>::TASKA_quant(const GTensor*, hQUANT, cudaStream_t stream_, int) : stream(stream_) {}
template <typename Typ>
TASKA_quant<Typ>::TASKA_quant(const GTensor*, int, int, bool, cudaStream_t stream_, int) : stream(stream_) {}
template struct TASKA_quant<floatX>;
