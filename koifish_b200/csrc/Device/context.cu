// context.cu -- device context, memory, CUDA-graph capture, NCCL plumbing for the C ABI in include/kf_device.h.
// Replaces InitCUDA / main_stream / gBUFF (reference: src/Device/CUDA/QKV.cu:501-571, huTensor.cu:70-103, 922-1003).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "kf_common.cuh"

struct kf_graph {
    cudaGraph_t graph    = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches    = 0;  // kernels recorded in the graph (added to ctx->launches at every replay)
};

extern "C" const char* kf_status_string(int s) {
    switch (s) {
        case KF_OK: return "KF_OK";
        case KF_ERR_NO_DEVICE: return "KF_ERR_NO_DEVICE: no CUDA device (there is no CPU fallback)";
        case KF_ERR_CUDA: return "KF_ERR_CUDA";
        case KF_ERR_BAD_ARG: return "KF_ERR_BAD_ARG";
        case KF_ERR_UNSUPPORTED: return "KF_ERR_UNSUPPORTED";
        case KF_ERR_OOM: return "KF_ERR_OOM";
        case KF_ERR_NCCL: return "KF_ERR_NCCL";
        case KF_ERR_QUANT: return "KF_ERR_QUANT";
    }
    return "KF_ERR_UNKNOWN";
}

extern "C" int kf_ctx_create(int device, void* cuda_stream, kf_ctx** out) {
    if (!out)
        return KF_ERR_BAD_ARG;
    *out  = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return KF_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n)
        return KF_ERR_BAD_ARG;
    kf_ctx* ctx = new kf_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return KF_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return KF_ERR_CUDA;
    }
    if (prop.major != 10) {
        // sm_100a cubins only: fail loudly instead of a silent fallback
        delete ctx;
        return KF_ERR_UNSUPPORTED;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return KF_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    *out = ctx;
    return KF_OK;
}

typedef ncclResult_t (*fn_ncclCommDestroy)(ncclComm_t);
static void* g_nccl = nullptr;
static void* nccl_sym(const char* name) {
    if (!g_nccl) {
        const char* cands[] = {"libnccl.so.2", "libnccl.so",
                               "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2", nullptr};
        for (int i = 0; cands[i] && !g_nccl; i++) g_nccl = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    }
    return g_nccl ? dlsym(g_nccl, name) : nullptr;
}

extern "C" int kf_ctx_destroy(kf_ctx* ctx) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    kf_p2p_destroy(ctx);
    kf_gemv_tma_destroy(ctx);
    kf_tmap_cache_destroy(ctx);
    if (ctx->nccl) {
        auto f = (fn_ncclCommDestroy)nccl_sym("ncclCommDestroy");
        if (f)
            f((ncclComm_t)ctx->nccl);
    }
    if (ctx->gemv_ws)
        cudaFree(ctx->gemv_ws);
    if (ctx->gemv_cnt)
        cudaFree(ctx->gemv_cnt);
    if (ctx->attn_ws)
        cudaFree(ctx->attn_ws);
    if (ctx->attn_cnt)
        cudaFree(ctx->attn_cnt);
    void* scratch[7] = {ctx->xperm, ctx->xnorm, ctx->tmp0, ctx->tmp1, ctx->deq_w, ctx->xg_buf, ctx->awq_ws};
    for (void* b : scratch)
        if (b)
            cudaFree(b);
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KF_OK;
}
extern "C" int kf_ctx_make_current(kf_ctx* ctx) {
    if (!ctx) return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    return KF_OK;
}
extern "C" int kf_ctx_sync(kf_ctx* ctx) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KF_OK;
}
extern "C" void* kf_ctx_stream(kf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int kf_ctx_sm_count(kf_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" const char* kf_last_error(kf_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
extern "C" uint64_t kf_launch_count(kf_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t kf_scratch_generation(kf_ctx* ctx) { return ctx ? ctx->scratch_gen : 0; }
extern "C" int kf_ctx_get_int(kf_ctx* ctx, const char* key, int* value_out) {
    if (!ctx || !key || !value_out) return KF_ERR_BAD_ARG;
    if (!strcmp(key, "gqa_min_ctx"))
        *value_out = ctx->gqa_min_ctx;
    else if (!strcmp(key, "tc_min_m"))
        *value_out = ctx->tc_min_m;
    else if (!strcmp(key, "pdl"))
        *value_out = ctx->pdl;
    else if (!strcmp(key, "attn_split"))
        *value_out = ctx->attn_split;
    else if (!strcmp(key, "deq_fma"))
        *value_out = ctx->deq_fma;
    else if (!strcmp(key, "gemv_exact"))
        *value_out = ctx->gemv_exact;
    else if (!strcmp(key, "gemv_last_s"))
        *value_out = ctx->gemv_last_s;
    else if (!strcmp(key, "tp_fused"))
        *value_out = ctx->tp_fused;
    else if (!strcmp(key, "gemv_xg_min_m"))
        *value_out = ctx->gemv_xg_min_m;
    else
        return KF_ERR_BAD_ARG;
    return KF_OK;
}
extern "C" int kf_ctx_set_int(kf_ctx* ctx, const char* key, int value) {
    if (!ctx || !key)
        return KF_ERR_BAD_ARG;
    if (!strcmp(key, "gemv_splitk"))
        ctx->gemv_splitk = value;
    else if (!strcmp(key, "gemv_variant"))
        ctx->gemv_variant = value;
    else if (!strcmp(key, "gemv_cluster"))
        ctx->gemv_cluster = value;
    else if (!strcmp(key, "gemv_exact"))
        ctx->gemv_exact = value;
    else if (!strcmp(key, "tp_fused"))
        ctx->tp_fused = value;
    else if (!strcmp(key, "gemv_xg_min_m"))
        ctx->gemv_xg_min_m = value;
    else if (!strcmp(key, "deq_fma"))
        ctx->deq_fma = value ? 1 : 0;
    else if (!strcmp(key, "pdl"))
        ctx->pdl = value;
    else if (!strcmp(key, "tc_min_m"))
        ctx->tc_min_m = value;
    else if (!strcmp(key, "attn_split"))
        ctx->attn_split = value;
    else if (!strcmp(key, "gqa_min_ctx"))
        ctx->gqa_min_ctx = value;
    else if (!strcmp(key, "attn_warps"))
        ctx->attn_warps = value;
    else if (!strcmp(key, "gemv_tma"))
        ctx->gemv_tma_on = value ? 1 : 0;
    else if (!strcmp(key, "gemv_tma_occ"))
        ctx->gemv_tma_occ = value;
    else if (!strcmp(key, "gemv_tma_smem_kb"))
        ctx->gemv_tma_smem_kb = value;
    else if (!strcmp(key, "gemv_tma_warps"))
        ctx->gemv_tma_warps = value;
#ifdef KF_DEBUG_KNOBS
    else if (!strcmp(key, "debug_skip"))
        ctx->debug_skip = value;
#endif
    else
        return KF_ERR_BAD_ARG;
    return KF_OK;
}

int kf_ensure_gemv_ws(kf_ctx* ctx, size_t bytes, int counters) {
    if (bytes > ctx->gemv_ws_bytes) {
        KF_REQUIRE(ctx, !ctx->capturing, "split-K workspace must be sized before graph capture (run one eager step first)");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->gemv_ws)
            cudaFree(ctx->gemv_ws);
        ctx->gemv_ws = nullptr, ctx->gemv_ws_bytes = 0;
        size_t want = bytes + bytes / 4;
        KF_CUDA(ctx, cudaMalloc(&ctx->gemv_ws, want));
        ctx->gemv_ws_bytes = want;
        ctx->scratch_gen++;
    }
    if (counters > ctx->gemv_cnt_n) {
        KF_REQUIRE(ctx, !ctx->capturing, "split-K counters must be sized before graph capture");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->gemv_cnt)
            cudaFree(ctx->gemv_cnt);
        ctx->gemv_cnt = nullptr, ctx->gemv_cnt_n = 0;
        int want = counters * 2 + 1024;
        KF_CUDA(ctx, cudaMalloc(&ctx->gemv_cnt, sizeof(unsigned) * want));
        KF_CUDA(ctx, cudaMemsetAsync(ctx->gemv_cnt, 0, sizeof(unsigned) * want, ctx->stream));
        ctx->gemv_cnt_n = want;
        ctx->scratch_gen++;
    }
    return KF_OK;
}
int kf_ensure_attn_ws(kf_ctx* ctx, size_t bytes) {
    if (bytes > ctx->attn_ws_bytes) {
        KF_REQUIRE(ctx, !ctx->capturing, "attention workspace must be sized before graph capture");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->attn_ws)
            cudaFree(ctx->attn_ws);
        ctx->attn_ws = nullptr, ctx->attn_ws_bytes = 0;
        KF_CUDA(ctx, cudaMalloc(&ctx->attn_ws, bytes * 2));
        ctx->attn_ws_bytes = bytes * 2;
        ctx->scratch_gen++;
    }
    return KF_OK;
}

int kf_ensure_buf(kf_ctx* ctx, void** buf, size_t* cap, size_t bytes) {
    if (bytes > *cap) {
        KF_REQUIRE(ctx, !ctx->capturing, "scratch buffers must be sized before graph capture (run one eager step first)");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*buf)
            cudaFree(*buf);
        *buf = nullptr, *cap = 0;
        KF_CUDA(ctx, cudaMalloc(buf, bytes + bytes / 8 + 256));
        *cap = bytes + bytes / 8;
        ctx->scratch_gen++;
    }
    return KF_OK;
}
int kf_ensure_attn_cnt(kf_ctx* ctx, int counters) {
    if (counters > ctx->attn_cnt_n) {
        KF_REQUIRE(ctx, !ctx->capturing, "attention counters must be sized before graph capture");
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->attn_cnt)
            cudaFree(ctx->attn_cnt);
        ctx->attn_cnt = nullptr, ctx->attn_cnt_n = 0;
        const int want = counters * 2 + 256;
        KF_CUDA(ctx, cudaMalloc(&ctx->attn_cnt, sizeof(unsigned) * want));
        KF_CUDA(ctx, cudaMemsetAsync(ctx->attn_cnt, 0, sizeof(unsigned) * want, ctx->stream));
        ctx->attn_cnt_n = want;
        ctx->scratch_gen++;
    }
    return KF_OK;
}

// ---------------------------------------------------------------- memory
extern "C" int kf_malloc(kf_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out)
        return KF_ERR_BAD_ARG;
    *out = nullptr;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 16));
    return KF_OK;
}
extern "C" int kf_free(kf_ctx* ctx, void* p) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    if (p)
        KF_CUDA(ctx, cudaFree(p));
    return KF_OK;
}
extern "C" int kf_memset(kf_ctx* ctx, void* p, int v, size_t bytes) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaMemsetAsync(p, v, bytes, ctx->stream));
    return KF_OK;
}
extern "C" int kf_h2d(kf_ctx* ctx, void* dev, const void* host, size_t bytes) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KF_OK;
}
extern "C" int kf_d2h(kf_ctx* ctx, void* host, const void* dev, size_t bytes) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return KF_OK;
}
extern "C" int kf_d2d(kf_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return KF_OK;
}
extern "C" int kf_host_alloc(size_t bytes, void** out) {
    if (!out)
        return KF_ERR_BAD_ARG;
    if (cudaMallocHost(out, bytes ? bytes : 16) != cudaSuccess) {
        (void)cudaGetLastError();
        return KF_ERR_OOM;
    }
    return KF_OK;
}
extern "C" int kf_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess)
        return KF_ERR_CUDA;
    return KF_OK;
}

// ---------------------------------------------------------------- CUDA graphs
extern "C" int kf_graph_begin(kf_ctx* ctx) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, !ctx->capturing, "already capturing");
    KF_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->capture_base = ctx->launches;
    return KF_OK;
}
extern "C" int kf_graph_end(kf_ctx* ctx, kf_graph** out) {
    if (!ctx || !out)
        return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, ctx->capturing, "not capturing");
    ctx->capturing = false;
    kf_graph* g    = new kf_graph();
    cudaError_t e  = cudaStreamEndCapture(ctx->stream, &g->graph);
    if (e != cudaSuccess) {
        ctx->last_error = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e);
        delete g;
        return KF_ERR_CUDA;
    }
    g->launches   = ctx->launches - ctx->capture_base;
    ctx->launches = ctx->capture_base;  // nothing ran during capture
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        ctx->last_error = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e);
        cudaGraphDestroy(g->graph);
        delete g;
        return KF_ERR_CUDA;
    }
    *out = g;
    return KF_OK;
}
extern "C" int kf_graph_launch(kf_ctx* ctx, kf_graph* g) {
    if (!ctx || !g)
        return KF_ERR_BAD_ARG;
    KF_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    return KF_OK;
}
extern "C" int kf_graph_destroy(kf_graph* g) {
    if (!g)
        return KF_ERR_BAD_ARG;
    if (g->exec)
        cudaGraphExecDestroy(g->exec);
    if (g->graph)
        cudaGraphDestroy(g->graph);
    delete g;
    return KF_OK;
}

// ---------------------------------------------------------------- NCCL (tensor parallel; loaded lazily so that a
// single-GPU process never needs libnccl)
typedef ncclResult_t (*fn_ncclGetUniqueId)(ncclUniqueId*);
typedef ncclResult_t (*fn_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_ncclAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_ncclAllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);

extern "C" int kf_nccl_unique_id(void* id_out) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    auto f = (fn_ncclGetUniqueId)nccl_sym("ncclGetUniqueId");
    if (!f || !id_out)
        return KF_ERR_NCCL;
    return f((ncclUniqueId*)id_out) == ncclSuccess ? KF_OK : KF_ERR_NCCL;
}
extern "C" int kf_ctx_init_nccl(kf_ctx* ctx, const void* id, int rank, int world) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world)
        return KF_ERR_BAD_ARG;
    ctx->rank = rank, ctx->world = world;
    if (world == 1)
        return KF_OK;
    auto f = (fn_ncclCommInitRank)nccl_sym("ncclCommInitRank");
    if (!f)
        return KF_ERR_NCCL;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclComm_t comm;
    if (f(&comm, world, uid, rank) != ncclSuccess)
        return KF_ERR_NCCL;
    ctx->nccl = (ncclComm*)comm;
    return KF_OK;
}
extern "C" int kf_allreduce_bf16(kf_ctx* ctx, void* buf, size_t count) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    if (ctx->world == 1)
        return KF_OK;
    auto f = (fn_ncclAllReduce)nccl_sym("ncclAllReduce");
    if (!f || !ctx->nccl)
        return KF_ERR_NCCL;
    if (f(buf, buf, count, ncclBfloat16, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
        return KF_ERR_NCCL;
    ctx->launches++;
    return KF_OK;
}
extern "C" int kf_allreduce_f32(kf_ctx* ctx, float* buf, size_t count) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    if (ctx->world == 1)
        return KF_OK;
    auto f = (fn_ncclAllReduce)nccl_sym("ncclAllReduce");
    if (!f || !ctx->nccl)
        return KF_ERR_NCCL;
    if (f(buf, buf, count, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
        return KF_ERR_NCCL;
    ctx->launches++;
    return KF_OK;
}
extern "C" int kf_allgather(kf_ctx* ctx, void* out, const void* in, size_t bytes_per_rank) {
    if (!ctx)
        return KF_ERR_BAD_ARG;
    if (ctx->world == 1) {
        if (out != in)
            KF_CUDA(ctx, cudaMemcpyAsync(out, in, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return KF_OK;
    }
    auto f = (fn_ncclAllGather)nccl_sym("ncclAllGather");
    if (!f || !ctx->nccl)
        return KF_ERR_NCCL;
    if (f(in, out, bytes_per_rank, ncclInt8, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
        return KF_ERR_NCCL;
    ctx->launches++;
    return KF_OK;
}
