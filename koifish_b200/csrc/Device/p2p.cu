// p2p.cu -- the tensor-parallel exchange step of decode over NVLink / NVSwitch PEER MEMORY, fused with the residual add.
//
// After the row-parallel matmuls (O and down) every rank holds an fp32 partial of the M x hidden activations; the reference has no
// counterpart (its multi_gpu.cuh is dead code).  The library path (ncclAllReduce + a residual-add kernel) costs
// two launches and ~15 us of latency per exchange -- 128 exchanges per Qwen3-32B token.  Here ONE kernel per exchange does
//     push : every rank stores its partial into slot [rank] of EVERY peer's symmetric buffer (remote stores over NVLink),
//     flag : publishes an epoch flag on every peer (release at system scope),
//     wait : spins on its own flags until all ranks' epochs arrived (acquire),
//     sum  : adds the slots in RANK ORDER (every rank gets bit-identical results) + the residual -> bf16 activations.
// One-shot all-reduce: latency = one NVLink store + one flag, the right trade for 20 KB messages.  Buffers alternate by epoch parity
// so a fast rank can start exchange e+1 while a slow rank still reads e (a rank reads e-1 before it signals e, and everybody waited
// for all signals of e before touching parity(e+1) = parity(e-1)).  The epoch lives in device memory and is advanced by the kernel,
// so the launch is CUDA-graph replayable (one counter per CTA index: a CTA only ever reads its own).  One process per GPU: the symmetric buffers are exchanged as CUDA IPC handles by the caller.
#include <string.h>

#include "kf_common.cuh"
#include "kf_tp.cuh"

namespace {
constexpr int kMaxWorld = 8;
constexpr int kP2PThreads = 256;
constexpr int kP2PCtaFloats = kP2PThreads * 4;  // one float4 per thread
constexpr int kMaxCtas = 512;

struct P2PState {
    int world = 0, rank = 0;
    size_t nmax = 0;              // floats per slot
    uint8_t* local = nullptr;     // this rank's symmetric buffer
    uint8_t* peer[kMaxWorld] = {};  // every rank's buffer mapped into this process (peer[rank] == local)
    unsigned* epoch = nullptr;    // device counters, one per CTA index
    size_t ll_off = 0;            // byte offset of the flag-in-data area of the fused exchange (kf_tp.cuh) inside a symmetric buffer
    unsigned* tok = nullptr;      // device token counter of the fused exchange
};
// layout of a symmetric buffer: flags[2][kMaxCtas][kMaxWorld] u32 | data[2][world][nmax] f32
__host__ __device__ inline size_t p2p_flags_bytes() { return (size_t)2 * kMaxCtas * kMaxWorld * 4; }

struct P2PParams {
    uint8_t* peer[kMaxWorld];
    unsigned* epoch;
    const float* partial;
    const uint16_t* residual;
    uint16_t* out;
    size_t n, nmax;
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kP2PThreads) kf_allreduce_residual_kernel(const P2PParams p) {
    const unsigned e      = p.epoch[blockIdx.x] + 1;  // per-CTA exchange counter: every rank counts the exchanges of chunk b identically
    const unsigned parity = e & 1;
    const size_t i0 = ((size_t)blockIdx.x * kP2PThreads + threadIdx.x) * 4;
    const bool active = i0 < p.n;
    kf_grid_dependency_wait();  // the partial comes from the matmul right before us
    // ---- push my partial into slot [rank] of every peer ----
    if (active) {
        const float4 v = *reinterpret_cast<const float4*>(p.partial + i0);
        for (int w = 0; w < p.world; w++) {
            const int peer = (p.rank + w) % p.world;  // spread the NVLink traffic: everybody starts at a different peer
            float* slot    = reinterpret_cast<float*>(p.peer[peer] + p2p_flags_bytes()) + ((size_t)parity * p.world + p.rank) * p.nmax;
            *reinterpret_cast<float4*>(slot + i0) = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    // ---- publish / await the epoch of this CTA's chunk ----
    if (threadIdx.x < p.world) {
        unsigned* remote = reinterpret_cast<unsigned*>(p.peer[threadIdx.x]) + ((size_t)parity * kMaxCtas + blockIdx.x) * kMaxWorld + p.rank;
        st_release_sys(remote, e);
        const unsigned* mine = reinterpret_cast<const unsigned*>(p.peer[p.rank]) + ((size_t)parity * kMaxCtas + blockIdx.x) * kMaxWorld + threadIdx.x;
        // a lost peer fails the launch instead of hanging the GPU; the bound is wall-clock (ranks may be seconds apart at the first
        // exchange, e.g. while one of them still quantises its shard)
        unsigned spins = 0;
        unsigned long long t_start = 0;
        while ((int)(ld_acquire_sys(mine) - e) < 0) {
            if ((++spins & 0x3ff) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (!t_start) t_start = now;
                if (now - t_start > 120ull * 1000000000ull) __trap();
            }
        }
    }
    __syncthreads();
    // ---- ordered sum + residual ----
    if (active) {
        const float* base = reinterpret_cast<const float*>(p.peer[p.rank] + p2p_flags_bytes()) + (size_t)parity * p.world * p.nmax;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < p.world; w++) {
            const float4 v = ld_sys_f4(reinterpret_cast<const float4*>(base + (size_t)w * p.nmax + i0));
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        const uint2 r = *reinterpret_cast<const uint2*>(p.residual + i0);
        uint2 o;
        // the same two roundings as the single-GPU epilogue: bf16(matmul), then bf16(residual + that)
        o.x = pack_bf16x2(bf16lo(r.x) + bf16_bits_to_f32(f32_to_bf16_bits(acc.x)), bf16hi(r.x) + bf16_bits_to_f32(f32_to_bf16_bits(acc.y)));
        o.y = pack_bf16x2(bf16lo(r.y) + bf16_bits_to_f32(f32_to_bf16_bits(acc.z)), bf16hi(r.y) + bf16_bits_to_f32(f32_to_bf16_bits(acc.w)));
        *reinterpret_cast<uint2*>(p.out + i0) = o;
    }
    __syncthreads();
    if (threadIdx.x == 0) p.epoch[blockIdx.x] = e;  // stream order: the next exchange starts after this kernel ends
}

P2PState* state_of(kf_ctx* ctx) { return reinterpret_cast<P2PState*>(ctx->p2p); }
}  // namespace

void kf_p2p_destroy(kf_ctx* ctx) {
    P2PState* s = state_of(ctx);
    if (!s) return;
    for (int w = 0; w < s->world; w++)
        if (s->peer[w] && w != s->rank) cudaIpcCloseMemHandle(s->peer[w]);
    if (s->local) cudaFree(s->local);
    if (s->epoch) cudaFree(s->epoch);
    delete s;
    ctx->p2p = nullptr;
}

extern "C" int kf_p2p_alloc(kf_ctx* ctx, size_t max_floats, int world, void* handle_out_64_bytes) {
    if (!ctx || !handle_out_64_bytes) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, world >= 2 && world <= kMaxWorld && max_floats >= 4, "2 <= world <= 8");
    KF_REQUIRE(ctx, (max_floats + kP2PCtaFloats - 1) / kP2PCtaFloats <= kMaxCtas, "message too large for the one-shot exchange");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    kf_p2p_destroy(ctx);
    P2PState* s = new P2PState();
    s->world = world, s->nmax = (max_floats + 3) & ~(size_t)3;
    s->ll_off          = (p2p_flags_bytes() + (size_t)2 * world * s->nmax * 4 + 255) & ~(size_t)255;
    const size_t bytes = s->ll_off + kf_tp_ll_bytes(world);
    KF_CUDA(ctx, cudaMalloc(&s->local, bytes));
    KF_CUDA(ctx, cudaMemset(s->local, 0, bytes));  // epoch 0 everywhere: the first exchange is epoch 1
    KF_CUDA(ctx, cudaMalloc(&s->epoch, kMaxCtas * 4 + 4));
    KF_CUDA(ctx, cudaMemset(s->epoch, 0, kMaxCtas * 4 + 4));
    s->tok = s->epoch + kMaxCtas;
    cudaIpcMemHandle_t h;
    KF_CUDA(ctx, cudaIpcGetMemHandle(&h, s->local));
    memcpy(handle_out_64_bytes, &h, 64);
    ctx->p2p = s;
    return KF_OK;
}

extern "C" int kf_p2p_attach(kf_ctx* ctx, const void* handles_world_x_64_bytes, int rank, int world) {
    if (!ctx || !handles_world_x_64_bytes) return KF_ERR_BAD_ARG;
    P2PState* s = state_of(ctx);
    KF_REQUIRE(ctx, s && s->world == world && rank >= 0 && rank < world, "kf_p2p_alloc first, same world");
    s->rank = rank;
    for (int w = 0; w < world; w++) {
        if (w == rank) {
            s->peer[w] = s->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles_world_x_64_bytes + (size_t)w * 64, 64);
        void* p = nullptr;
        KF_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer[w] = (uint8_t*)p;
    }
    KF_CUDA(ctx, cudaDeviceSynchronize());
    return KF_OK;
}

extern "C" int kf_p2p_release(kf_ctx* ctx) {
    if (!ctx) return KF_ERR_BAD_ARG;
    cudaStreamSynchronize(ctx->stream);
    kf_p2p_destroy(ctx);
    return KF_OK;
}
// ---- the fused exchange (kf_tp.cuh): view for the matmul kernels, token counter, unpack ------------------------------------------
int kf_tp_view(kf_ctx* ctx, KfTpView* out) {
    P2PState* s = state_of(ctx);
    if (!s || !s->peer[0] || !ctx->tp_fused) return KF_ERR_UNSUPPORTED;
    memset(out, 0, sizeof(*out));
    for (int w = 0; w < s->world; w++) out->peer[w] = s->peer[w] + s->ll_off;
    out->tok = s->tok, out->world = s->world, out->rank = s->rank, out->stride = ctx->tp_stride;
    return KF_OK;
}
namespace {
__global__ void kf_tp_begin_kernel(unsigned* tok) {
    kf_grid_dependency_wait();  // the previous forward (its last consumer) is complete
    *tok = *tok + 1;
}
}  // namespace
extern "C" int kf_tp_begin(kf_ctx* ctx) {
    if (!ctx) return KF_ERR_BAD_ARG;
    P2PState* s = state_of(ctx);
    ctx->tp_xid = 0;
    if (!s || !s->peer[0] || !ctx->tp_fused) return KF_OK;  // nothing to advance: the callers take the unfused path
    KF_CUDA(ctx, kf_launch_pdl(ctx, kf_tp_begin_kernel, dim3(1), dim3(1), 0, s->tok));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
extern "C" int kf_exchange_fused_ready(kf_ctx* ctx, int M, int cols) {
    KfTpView v;
    return ctx && M >= 1 && M <= 8 && cols % 128 == 0 && (size_t)M * cols <= KF_TP_LL_ELEMS && kf_tp_view(ctx, &v) == KF_OK ? 1 : 0;
}
extern "C" int kf_p2p_ready(kf_ctx* ctx) { return ctx && state_of(ctx) && state_of(ctx)->peer[0] ? 1 : 0; }

// out = residual + sum over ranks of partial (both roundings of the single-GPU path); out may alias residual
extern "C" int kf_allreduce_residual(kf_ctx* ctx, void* out_bf16, const void* residual_bf16, const float* partial, size_t n) {
    if (!ctx || !out_bf16 || !residual_bf16 || !partial) return KF_ERR_BAD_ARG;
    P2PState* s = state_of(ctx);
    if (!s || !s->peer[0] || n > s->nmax || (n & 3)) {  // library path: NCCL all-reduce on a copy-free in-place buffer + add
        int rc = kf_allreduce_f32(ctx, const_cast<float*>(partial), n);
        if (!rc) rc = kf_residual_add_f32(ctx, out_bf16, residual_bf16, partial, n);
        return rc;
    }
    P2PParams p;
    memset(&p, 0, sizeof(p));
    for (int w = 0; w < s->world; w++) p.peer[w] = s->peer[w];
    p.epoch = s->epoch, p.partial = partial, p.residual = (const uint16_t*)residual_bf16, p.out = (uint16_t*)out_bf16;
    p.n = n, p.nmax = s->nmax, p.world = s->world, p.rank = s->rank;
    const unsigned ctas = (unsigned)((n + kP2PCtaFloats - 1) / kP2PCtaFloats);
    KF_CUDA(ctx, kf_launch_pdl(ctx, kf_allreduce_residual_kernel, dim3(ctas), dim3(kP2PThreads), 0, p));
    KF_LAUNCH_CHECK(ctx);
    return KF_OK;
}
