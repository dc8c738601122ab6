// kf_common.cuh -- shared definitions of the device layer (context, error handling, bf16 / PTX helpers).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "kf_device.h"

struct ncclComm;

struct kf_ctx {
    static constexpr int kMaxDevices = 64;
    int device           = 0;
    cudaStream_t stream  = nullptr;
    bool own_stream      = false;
    int sm_count         = 148;
    uint64_t launches    = 0;
    bool capturing       = false;
    // split-K workspace (fp32 partials) + per-row-block arrival counters (self-resetting)
    float* gemv_ws       = nullptr;
    size_t gemv_ws_bytes = 0;
    unsigned* gemv_cnt   = nullptr;
    int gemv_cnt_n       = 0;
    // attention split workspace
    float* attn_ws       = nullptr;
    size_t attn_ws_bytes = 0;
    unsigned* attn_cnt   = nullptr;  // per (token, head) arrival counters of the fused attention (self-resetting)
    int attn_cnt_n       = 0;
    int pdl              = 1;        // programmatic dependent launch between consecutive kernels of a decode step
    // scratch of the tensor-core (M > 64) path: permuted / normalised activations and the gate / up panels of a SwiGLU
    void *xperm = nullptr, *xnorm = nullptr, *tmp0 = nullptr, *tmp1 = nullptr;
    size_t xperm_bytes = 0, xnorm_bytes = 0, tmp0_bytes = 0, tmp1_bytes = 0;
    void* xg_buf       = nullptr;  // activations pre-staged in fragment order for the 4..8-token decode GEMV (gemv.cu, XG variants)
    size_t xg_bytes    = 0;
    int gemv_xg_min_m  = 3;        // token count from which they are, where the k-split plan would otherwise need two waves (0 = never, < 0 = always from |n|)
    void* awq_ws        = nullptr;  // k-slice partial sums of the AWQ GEMV (awq.cu)
    size_t awq_ws_bytes = 0;
    void* deq_w        = nullptr;  // dequantised copy of ONE NormalFloat4 weight for the many-token path (nf4.cu)
    size_t deq_w_bytes = 0;
    int tc_min_m = -1;  // token count from which kf_linear* use the tcgen05 GEMM: -1 = per weight type (linear.cu), 0 = never
    // tuning
    int gemv_splitk  = 0;
    int gemv_variant = 0;
    int gemv_cluster = 1;  // M = 1 split-K: 1 = merge the k-slices of a row block inside a thread-block cluster (DSMEM) when S <= 8; 2 = also cap S at 8; 0 = global workspace
    // 0 (default): decode GEMVs over 4-bit weights apply the group's affine map to fp32 group sums of fp16 codes (MODE_FAST, gemv.cu: the ALU pipe
    // cannot unpack bit-exact bf16 weights at the HBM rate) -- logits within the stated tolerance ; 1: the in-kernel dequant reproduces the
    // reference's bf16 roundings bit for bit (weights identical to kf_dequant / GetDataX inside the matmul), ~0.57 of the HBM peak
    int gemv_exact   = 0;
    int gemv_last_s  = 0;  // read-only: k-split of the most recent dequant-GEMV launch (tools/gemv_bench.py records it)
    // rounding of the reference's dequant expression (step * k - zero) in bf16 (CU_Q128toX_, T.cu:274), pinned against the reference kernel
    // compiled both ways (oracle/ref_kernels.cu): 1 = ONE rounding, fma.rn.bf16 -- what nvcc's default -fmad=true (implied by the
    // reference's -use_fast_math) generates on sm_90+, i.e. what the reference computes on a B200 ; 0 = TWO roundings (bf16 multiply, then
    // bf16 subtract): -fmad=false and every pre-sm_90 build
    int deq_fma      = 1;
    int attn_split   = 0;
    int gqa_min_ctx  = 1024;  // single-sequence decode: contexts beyond this use the kv-group tensor-core attention (Transformer layer reads it)
    int attn_warps   = 0;  // warps per CTA of the cluster attention (0 = default)
    int debug_skip   = 0;  // only honoured in builds with -DKF_DEBUG_KNOBS (timing experiments: bit 0 skips attention, bit 1 the skinny GEMVs)
    // persistent TMA-fed stream-K GEMV (gemv_tma.cu): on / CTAs per SM / shared-memory budget per SM in KB
    int gemv_tma_on = 0, gemv_tma_occ = 1, gemv_tma_smem_kb = 112, gemv_tma_warps = 8;
    void* gemv_tma  = nullptr;  // its state (tensor-map cache, stream-K workspace)
    // bumped whenever a context scratch buffer is reallocated: CUDA graphs captured before hold stale pointers and must be re-captured
    uint64_t scratch_gen = 0;
    uint64_t capture_base = 0;  // launches counted when the current graph capture began
    // tensor parallel
    ncclComm* nccl = nullptr;
    int rank = 0, world = 1;
    void* tmap_cache = nullptr;  // encoded TMA tensor maps, keyed by (address, geometry, box) (gemm_tc.cu)
    void* p2p = nullptr;  // peer-memory exchange state (p2p.cu)
    int tp_fused  = 1;    // knob: the decode exchange rides on the matmul kernels (kf_tp.cuh) when the peer buffers are attached
    int tp_stride = 256;  // epochs reserved per forward (>= exchanges per forward, even)
    int tp_xid    = 0;    // ordinal of the next fused exchange of the current forward (kf_tp_begin resets it)
    std::string last_error;
};

#define KF_CUDA(ctx, expr)                                                                                          \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess) {                                                                                    \
            if (ctx) {                                                                                              \
                char _b[512];                                                                                       \
                snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
                (ctx)->last_error = _b;                                                                             \
            }                                                                                                       \
            return _e == cudaErrorMemoryAllocation ? KF_ERR_OOM : KF_ERR_CUDA;                                      \
        }                                                                                                           \
    } while (0)

#define KF_REQUIRE(ctx, cond, msg)                                                        \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            if (ctx) {                                                                    \
                char _b[512];                                                             \
                snprintf(_b, sizeof(_b), "%s:%d requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
                (ctx)->last_error = _b;                                                   \
            }                                                                             \
            return KF_ERR_BAD_ARG;                                                        \
        }                                                                                 \
    } while (0)

#define KF_LAUNCH_CHECK(ctx)                    \
    do {                                        \
        (ctx)->launches++;                      \
        KF_CUDA(ctx, cudaGetLastError());       \
    } while (0)

static inline int kf_type_bits(int type) {
    switch (type) {
        case KF_T_BF16: return 16;
        case KF_T_F8E5M2: return 8;
        case KF_T_Q4:
        case KF_T_NF4:
        case KF_T_AWQ4: return 4;
        case KF_T_Q2:
        case KF_T_SIGN: return 2;
        case KF_T_BINARY: return 1;
    }
    return 0;
}
static inline bool kf_type_packed(int type) { return type == KF_T_Q4 || type == KF_T_Q2 || type == KF_T_SIGN || type == KF_T_BINARY; }

// gama blob: [R_SCALE rows][C_SCALE cols][ZERO nG][STEP nG]  (src/Tensor/GTensor.cpp:456-510)
static inline const uint16_t* kf_gama_zero(const kf_tensor_desc& w) {
    if (w.zero_dev) return (const uint16_t*)w.zero_dev;
    return (const uint16_t*)w.gama_dev + w.rows + w.cols;
}
static inline const uint16_t* kf_gama_step(const kf_tensor_desc& w) {
    if (w.step_dev) return (const uint16_t*)w.step_dev;
    return (const uint16_t*)w.gama_dev + w.rows + w.cols + ((size_t)w.rows * w.cols) / w.group;
}
static inline bool kf_has_gama(const kf_tensor_desc& w) { return w.gama_dev || (w.zero_dev && w.step_dev); }

void kf_p2p_destroy(kf_ctx* ctx);
void kf_gemv_tma_destroy(kf_ctx* ctx);
void kf_tmap_cache_destroy(kf_ctx* ctx);  // gemm_tc.cu
// gemv_tma.cu: KF_OK = launched, 1 = request not covered by this kernel (caller falls back to gemv.cu), < 0 = error
int kf_gemv_tma(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                const void* norm_w, float norm_eps);
int kf_ensure_gemv_ws(kf_ctx* ctx, size_t bytes, int counters);
// awq.cu
int kf_awq_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out_bf16, int transposed);
int kf_awq_gemv(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual);
// nf4.cu
int kf_nf4_quantize(kf_ctx* ctx, const void* w_bf16, int rows, int cols, void* data, void* gama);
int kf_nf4_dequant(kf_ctx* ctx, const kf_tensor_desc* w, void* out_bf16);
int kf_nf4_embed(kf_ctx* ctx, void* out, const kf_tensor_desc* w, const int32_t* tokens, int M);
int kf_nf4_gemv(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual);
int kf_axb_epilogue(kf_ctx* ctx, void* d, const float* acc, const void* bias, float alpha, float beta, int rows, size_t n);  // ops.cu
int kf_ensure_attn_ws(kf_ctx* ctx, size_t bytes);
int kf_qknorm_rope_kv_warp(kf_ctx* ctx, void* q, const void* k, const void* v, const void* qw, const void* kw, void* kcache, void* vcache,
                           const void* table, const int32_t* pos_dev, int M, int n_head, int n_kv, int hd, float eps, size_t seq_stride);
int kf_ensure_attn_cnt(kf_ctx* ctx, int counters);
int kf_ensure_buf(kf_ctx* ctx, void** buf, size_t* cap, size_t bytes);
// gemm_tc.cu: tcgen05 / TMEM dequant-GEMM; epilogue 0 none / 1 residual / 4 fp32.  xp = the activations in the k order the kernel
// wants for w's type, as returned by kf_tc_prepare_x (x itself, or the context scratch holding the permuted copy)
int kf_tc_prepare_x(kf_ctx* ctx, const kf_tensor_desc* w, const void* x, int M, const void** xp_out);
int kf_tc_prepare_x_norm(kf_ctx* ctx, const kf_tensor_desc* w, const void* x, const void* norm_w, float eps, int M, const void** xp_out);
int kf_tc_same_order(const kf_tensor_desc* a, const kf_tensor_desc* b);  // 0: one prepared copy serves both weights
int kf_gemm_tc(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* xp, int M, int epilogue, const void* residual);
// up to 3 weights of the same type / K / group in ONE launch (Q/K/V, gate/up)
int kf_gemm_tc_multi(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* xp, int M, int epilogue, const void* residual);

#ifdef __CUDACC__
// Launch with the programmatic-dependent-launch attribute (when ctx->pdl): the kernel may start while its predecessor in the stream
// is still draining; it must execute kf_grid_dependency_wait() before touching anything the predecessor writes.
template <typename... KArgs, typename... Args>
static inline cudaError_t kf_launch_pdl(kf_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = ctx->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// thread-block cluster barrier (split arrive / wait) and a store into the shared memory of CTA `cta_rank` of the cluster
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(float* local_ptr, int cta_rank, float v) {
    uint32_t remote;
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr);
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta_rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
__device__ __forceinline__ void kf_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kf_grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
// gemv.cu: skinny path (M <= 64); epilogue 0 none / 1 residual / 2 swiglu(gate = w[0], up = w[1])
// norm_w != nullptr: x is RMS-normalised (weights norm_w, eps norm_eps) while it is staged, as kf_rmsnorm would have done
int kf_gemv_small(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                  const void* norm_w, float norm_eps);

#ifdef __CUDACC__
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ __nv_bfloat162 u32_as_bf162(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
// the reference's dequant of one code: (step * (bf16)k - zero) with bf16 operators (CU_Q128toX_, src/Device/CUDA/T.cu:274).  FMA = true:
// contracted to one fma.rn.bf16 (the reference built with nvcc defaults / -use_fast_math for sm_90+) ; false: multiply and subtract each
// round (-fmad=false, pre-sm_90).  The _rn intrinsics are never re-contracted by ptxas.
template <bool FMA>
__device__ __forceinline__ uint16_t kf_deq_scalar(int k, uint16_t step, uint16_t zero) {
    const __nv_bfloat16 kk = __int2bfloat16_rn(k), s = __ushort_as_bfloat16(step), z = __ushort_as_bfloat16(zero);
    if (FMA) return __bfloat16_as_ushort(__hfma(s, kk, __hneg(z)));
    return __bfloat16_as_ushort(__hsub_rn(__hmul_rn(s, kk), z));
}
__device__ __forceinline__ uint32_t bf162_as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// streaming 128-bit load: weights are read exactly once per token -> do not allocate in L1
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum with a fixed reduction order (warp butterflies, then the warps in index order): deterministic, and shared by the
// stand-alone RMSNorm kernel and the RMSNorm folded into the GEMV so that both produce the same scale bit for bit.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif
