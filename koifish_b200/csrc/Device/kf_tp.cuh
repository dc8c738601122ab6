// kf_tp.cuh -- the tensor-parallel decode exchange fused INTO the matmul kernels (gemv.cu): no launch of its own.
//
// After a row-parallel matmul (O, down) every rank holds an fp32 partial of the [M x hidden] activations.  The stand-alone kernel of
// p2p.cu (kf_allreduce_residual) costs a launch boundary, a fence round trip and a flag per exchange.  Here the exchange rides on the
// kernels either side of it, in the flag-in-data ("LL") style: every 8 bytes that cross NVLink carry 4 bytes of payload and the 4-byte
// epoch of the exchange, so that a reader polls the data itself -- one NVLink traversal, no fence, no separate flag.
//
//   epilogue of the O / down GEMV, the CTA that holds the final fp32 tile of row block rb, on EVERY rank:
//     push   : store the tile into slot [rank] of every peer -- {v0, e, v1, e} 16-byte stores, each 8-byte half self-validating;
//     reduce : poll the `world - 1` remote slots of that row block in local memory, add the partials in RANK ORDER (bit-identical on
//              every rank), round to bf16, add the residual, round (the two roundings of the single-GPU epilogue), write the rows of y.
//   The next kernel of the stream (RMSNorm + QKV / gate-up) reads plain activations after its griddepcontrol.wait and knows nothing of
//   the exchange: one NVLink traversal per exchange, no launch, no fence, no flag, no extra pass over the activations.
//
// Epoch of an exchange = token counter * stride + ordinal + 1: the token counter sits in device memory and is advanced by kf_tp_begin
// (one thread, first kernel of a forward), the ordinal is baked into the launch -- CUDA-graph replayable.  Two parities suffice: a peer
// writes exchange e + 2 into the slots of e only after its epilogue of e + 1 completed, which needed OUR partial of e + 1, which our
// stream issues after every CTA of our epilogue of e has finished reading.
#pragma once
#include "kf_common.cuh"

constexpr int KF_TP_MAX_WORLD   = 8;
constexpr size_t KF_TP_LL_ELEMS = 8 * 8192;  // elements per LL slot: M <= 8 tokens x hidden <= 8192

struct KfTpView {                     // by-value kernel parameter
    uint8_t* peer[KF_TP_MAX_WORLD];  // LL area of every rank's symmetric buffer, mapped here (peer[rank] is local)
    const unsigned* tok;              // device token counter
    int world, rank, stride;          // stride: epochs reserved per token (>= exchanges per token, even)
};
// LL area: scat[2 parities][world][LL_ELEMS] x {f32, epoch}
__host__ __device__ inline size_t kf_tp_ll_bytes(int world) { return (size_t)2 * world * KF_TP_LL_ELEMS * 8; }
// p2p.cu
int kf_tp_view(kf_ctx* ctx, KfTpView* out);  // KF_OK when the fused exchange is available on this context

#ifdef __CUDACC__
namespace kftp {
__device__ __forceinline__ void st16_sys(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld16_sys(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// call AFTER kf_grid_dependency_wait(): the counter is advanced by the first kernel of the forward
__device__ __forceinline__ unsigned epoch(const KfTpView& v, int xid) {
    return *reinterpret_cast<const volatile unsigned*>(v.tok) * (unsigned)v.stride + (unsigned)xid + 1u;
}
__device__ __forceinline__ uint8_t* scat(const KfTpView& v, int dst_rank, unsigned parity, int src_rank) {
    return v.peer[dst_rank] + ((size_t)(parity * v.world + src_rank) * KF_TP_LL_ELEMS) * 8;
}
// a lost peer fails the launch instead of hanging the GPU: wall-clock bound on every poll loop (ranks may be seconds apart at the first
// exchange, e.g. while one of them still quantises its shard)
struct SpinGuard {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void tick() {
        if ((++spins & 0x3ff) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            if (now - t0 > 120ull * 1000000000ull) __trap();
        }
    }
};
}  // namespace kftp
#endif
