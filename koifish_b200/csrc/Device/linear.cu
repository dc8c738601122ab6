// linear.cu -- the matmul entry points of the C ABI (TASKA_AxB::blasLt as used by SLP::Forw; reference
// src/Tensor/GTensor.hpp:703-741, src/Device/CUDA/NeuronFuse.cu:305-381).  M <= 64 tokens go to the HBM-bound fused
// dequant-GEMV (gemv.cu); larger M to the tcgen05 / TMEM dequant-GEMM (gemm_tc.cu).
#include "kf_common.cuh"

// same storage type, K, group and bias: the weights can share one launch
static bool tc_fusable(const kf_tensor_desc* a, const kf_tensor_desc* b) {
    return a->type == b->type && a->cols == b->cols && a->group == b->group && a->qbias == b->qbias;
}

// M > 64 through the skinny kernel in 64-token panels (used when the tensor-core path is switched off: ctx knob tc_min_m = 0)
static int linear_panels_same(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                              const void* norm_w, float norm_eps) {
    if (M <= 64) return kf_gemv_small(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
    const int K = w[0].cols;
    for (int m0 = 0; m0 < M; m0 += 64) {
        const int mm = M - m0 < 64 ? M - m0 : 64;
        void* yy[3];
        for (int i = 0; i < n; i++) {
            const int rows = (epilogue == 2 && i == 1) ? 0 : w[i].rows;
            yy[i]          = y[i] ? (void*)((char*)y[i] + (size_t)m0 * rows * (epilogue == KF_EPI_F32 ? 4 : 2)) : nullptr;
        }
        if (epilogue == 2) yy[1] = yy[0];
        const void* res = residual ? (const void*)((const uint16_t*)residual + (size_t)m0 * w[0].rows) : nullptr;
        int rc = kf_gemv_small(ctx, n, yy, w, (const uint16_t*)x + (size_t)m0 * K, mm, epilogue, res, norm_w, norm_eps);
        if (rc) return rc;
    }
    return KF_OK;
}
// The quantizer card selects the storage type per tensor-name substring (QUANT_CARD::Init4Neuron, reference src/Tensor/GeQuant.cpp:
// 1186-1285), so Q / K / V (or gate / up) of one block may differ in type, group or bias, e.g. {"q_proj": {"quant_method": "RTN", "bits": 4}}.
// Weights that agree share one launch; the others get their own (the RMSNorm folded into each: same arithmetic, same result).
static int linear_panels(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                         const void* norm_w, float norm_eps) {
    bool same = true;
    for (int i = 1; i < n; i++) same = same && tc_fusable(&w[0], &w[i]);
    if (same) return linear_panels_same(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
    if (epilogue == 2) {  // SwiGLU of a mixed gate / up pair: two matmuls into scratch, then CU_swiglu_v0 as a stand-alone op
        const size_t bytes = (size_t)M * w[0].rows * 2;
        int rc = kf_ensure_buf(ctx, &ctx->tmp0, &ctx->tmp0_bytes, bytes);
        if (!rc) rc = kf_ensure_buf(ctx, &ctx->tmp1, &ctx->tmp1_bytes, bytes);
        void* g1[1] = {ctx->tmp0};
        void* u1[1] = {ctx->tmp1};
        if (!rc) rc = linear_panels_same(ctx, 1, g1, &w[0], x, M, 0, nullptr, norm_w, norm_eps);
        if (!rc) rc = linear_panels_same(ctx, 1, u1, &w[1], x, M, 0, nullptr, norm_w, norm_eps);
        if (!rc) rc = kf_swiglu(ctx, y[0], ctx->tmp0, ctx->tmp1, (size_t)M * w[0].rows);
        return rc;
    }
    bool done[3] = {false, false, false};
    for (int i = 0; i < n; i++) {
        if (done[i]) continue;
        kf_tensor_desc gw[3];
        void* gy[3];
        int gn = 0;
        for (int j = i; j < n; j++)
            if (!done[j] && tc_fusable(&w[i], &w[j])) gw[gn] = w[j], gy[gn] = y[j], gn++, done[j] = true;
        int rc = linear_panels_same(ctx, gn, gy, gw, x, M, epilogue, residual, norm_w, norm_eps);
        if (rc) return rc;
    }
    return KF_OK;
}

// Token count from which the tensor-core kernel beats the skinny one (measured on B200): bf16 weights always (TMA streams them at the
// full HBM rate), everything else from 9 tokens.  Launch by launch the packed formats cross over between 12 and 16 tokens
// (profiles/r01_tc_crossover.txt), but inside the model the tensor-core path also shares one launch between Q/K/V and between gate/up,
// and the skinny kernel's 16-token variant is its weakest: Qwen3-32B batch 12 decodes at 972 tok/s this way against 689
// (profiles/r01_tc_crossover.txt, in-model table).  ctx knob tc_min_m: -1 auto, 0 never, n > 0 fixed threshold.
static bool use_tensor_cores(const kf_ctx* ctx, int n, const kf_tensor_desc* w, int M) {
    if (ctx->tc_min_m == 0) return false;
    for (int i = 0; i < n; i++) {
        if (w[i].cols % 128 != 0 || w[i].rows % 16 != 0 || ((uintptr_t)w[i].data_dev & 15)) return false;
        const int need = ctx->tc_min_m > 0 ? ctx->tc_min_m : w[i].type == KF_T_BF16 ? 1 : 9;
        if (M < need) return false;
    }
    return true;
}

// epilogue: 0 none, 1 residual, 2 swiglu(w[0] gate, w[1] up -> y[0]), 4 fp32
static int linear_any(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                      const void* norm_w, float norm_eps);
// NormalFloat4 (nf4.cu) and vendor-AWQ (awq.cu) weights: RMSNorm as its own launch, then per weight either the warp-per-row LUT GEMV (<= 8 tokens) or dequantise into a
// context scratch + the bf16 tensor-core GEMM (the reference's own route for every format: GTensor::GetDataX + cuBLASLt)
static int linear_nf4(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                      const void* norm_w, float norm_eps) {
    const int K = w[0].cols;
    int rc      = KF_OK;
    if (norm_w) {
        rc = kf_ensure_buf(ctx, &ctx->xnorm, &ctx->xnorm_bytes, (size_t)M * K * 2);
        if (!rc) rc = kf_rmsnorm(ctx, ctx->xnorm, x, norm_w, M, K, norm_eps);
        if (rc) return rc;
        x = ctx->xnorm;
    }
    auto one = [&](void* yy, const kf_tensor_desc& ww, int epi, const void* res) -> int {
        if (ww.type != KF_T_NF4 && ww.type != KF_T_AWQ4) {
            void* ys[1] = {yy};
            return linear_any(ctx, 1, ys, &ww, x, M, epi, res, nullptr, 0.f);
        }
        if (M <= 8) return ww.type == KF_T_NF4 ? kf_nf4_gemv(ctx, yy, &ww, x, M, epi, res) : kf_awq_gemv(ctx, yy, &ww, x, M, epi, res);
        int r = kf_ensure_buf(ctx, &ctx->deq_w, &ctx->deq_w_bytes, (size_t)ww.rows * ww.cols * 2);
        if (!r) r = ww.type == KF_T_NF4 ? kf_nf4_dequant(ctx, &ww, ctx->deq_w) : kf_awq_dequant(ctx, &ww, ctx->deq_w, 1);
        if (r) return r;
        kf_tensor_desc bw = {};
        bw.data_dev = ctx->deq_w, bw.rows = ww.rows, bw.cols = ww.cols, bw.type = KF_T_BF16;
        void* ys[1] = {yy};
        return linear_any(ctx, 1, ys, &bw, x, M, epi, res, nullptr, 0.f);
    };
    if (epilogue == 2) {  // SwiGLU: gate and up into scratch, then CU_swiglu_v0
        const size_t bytes = (size_t)M * w[0].rows * 2;
        rc = kf_ensure_buf(ctx, &ctx->tmp0, &ctx->tmp0_bytes, bytes);
        if (!rc) rc = kf_ensure_buf(ctx, &ctx->tmp1, &ctx->tmp1_bytes, bytes);
        if (!rc) rc = one(ctx->tmp0, w[0], 0, nullptr);
        if (!rc) rc = one(ctx->tmp1, w[1], 0, nullptr);
        if (!rc) rc = kf_swiglu(ctx, y[0], ctx->tmp0, ctx->tmp1, (size_t)M * w[0].rows);
        return rc;
    }
    for (int i = 0; i < n && !rc; i++) rc = one(y[i], w[i], epilogue, residual);
    return rc;
}
static int linear_any(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual,
                      const void* norm_w, float norm_eps) {
    for (int i = 0; i < n; i++)
        if (w[i].type == KF_T_NF4 || w[i].type == KF_T_AWQ4) return linear_nf4(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
    if (!use_tensor_cores(ctx, n, w, M)) return linear_panels(ctx, n, y, w, x, M, epilogue, residual, norm_w, norm_eps);
    // ---- tensor-core path: RMSNorm (if any) once into a scratch, then one tcgen05 GEMM per weight ----
    const int K    = w[0].cols;
    const void* xin = x;
    const void* xp  = nullptr;
    int rc          = KF_OK;
    bool one_order  = true;  // every weight wants the same activation order: RMSNorm can be written in that order directly
    for (int i = 1; i < n; i++) one_order = one_order && !kf_tc_same_order(&w[0], &w[i]);
    if (norm_w && one_order) {
        rc = kf_tc_prepare_x_norm(ctx, &w[0], x, norm_w, norm_eps, M, &xp);  // norm + permutation in one launch
        if (rc) return rc;
    } else {
        if (norm_w) {
            rc = kf_ensure_buf(ctx, &ctx->xnorm, &ctx->xnorm_bytes, (size_t)M * K * 2);
            if (!rc) rc = kf_rmsnorm(ctx, ctx->xnorm, x, norm_w, M, K, norm_eps);
            if (rc) return rc;
            xin = ctx->xnorm;
        }
        // the activations in the k order of the weights' type: prepared once, shared by every weight of the same type
        rc = kf_tc_prepare_x(ctx, &w[0], xin, M, &xp);
        if (rc) return rc;
    }
    if (epilogue == 2) {
        const size_t bytes = (size_t)M * w[0].rows * 2;
        rc = kf_ensure_buf(ctx, &ctx->tmp0, &ctx->tmp0_bytes, bytes);
        if (!rc) rc = kf_ensure_buf(ctx, &ctx->tmp1, &ctx->tmp1_bytes, bytes);
        if (!rc && tc_fusable(&w[0], &w[1])) {  // gate and up in one launch
            void* gu[2] = {ctx->tmp0, ctx->tmp1};
            rc          = kf_gemm_tc_multi(ctx, 2, gu, w, xp, M, 0, nullptr);
        } else {
            if (!rc) rc = kf_gemm_tc(ctx, ctx->tmp0, &w[0], xp, M, 0, nullptr);
            if (!rc && kf_tc_same_order(&w[0], &w[1])) rc = kf_tc_prepare_x(ctx, &w[1], xin, M, &xp);
            if (!rc) rc = kf_gemm_tc(ctx, ctx->tmp1, &w[1], xp, M, 0, nullptr);
        }
        if (!rc) rc = kf_swiglu(ctx, y[0], ctx->tmp0, ctx->tmp1, (size_t)M * w[0].rows);  // CU_swiglu_v0 on the bf16 gate / up, as the reference
        return rc;
    }
    bool fuse = n > 1 && epilogue != KF_EPI_RESIDUAL;
    for (int i = 1; i < n && fuse; i++) fuse = tc_fusable(&w[0], &w[i]);
    if (fuse) return kf_gemm_tc_multi(ctx, n, y, w, xp, M, epilogue, nullptr);  // Q / K / V in one launch
    for (int i = 0; i < n; i++) {
        if (i > 0 && kf_tc_same_order(&w[i - 1], &w[i])) rc = kf_tc_prepare_x(ctx, &w[i], xin, M, &xp);
        if (!rc) rc = kf_gemm_tc(ctx, y[i], &w[i], xp, M, epilogue, residual);
        if (rc) return rc;
    }
    return KF_OK;
}

extern "C" int kf_linear(kf_ctx* ctx, void* y, const kf_tensor_desc* w, const void* x, int M, int epilogue, const void* residual) {
    if (!ctx || !y || !w || !x) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, epilogue == KF_EPI_NONE || epilogue == KF_EPI_RESIDUAL || epilogue == KF_EPI_F32, "epilogue");
    KF_REQUIRE(ctx, M >= 1, "M");
    void* ys[1] = {y};
    return linear_any(ctx, 1, ys, w, x, M, epilogue, residual, nullptr, 0.f);
}
// TASKA_AxB in full (src/Tensor/GTensor.hpp:698-741): d = alpha * x . w^T + beta * d + bias, bias one bf16 per output row or NULL.  The
// matmul runs through the same kernels with fp32 partial output, then one fp32 epilogue rounds once -- as cuBLASLt's epilogue does.
// alpha = 1, beta = 0, bias = NULL is kf_linear(..., KF_EPI_NONE) (SLP::Forw of the inference path).
extern "C" int kf_linear_axb(kf_ctx* ctx, void* d, const kf_tensor_desc* w, const void* x, int M, float alpha, float beta, const void* bias) {
    if (!ctx || !d || !w || !x) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1, "M");
    if (alpha == 0.f && beta == 0.f && !bias) return KF_OK;  // TASKA_AxB::isPass
    if (alpha == 1.f && beta == 0.f && !bias) return kf_linear(ctx, d, w, x, M, KF_EPI_NONE, nullptr);
    const size_t n = (size_t)M * w->rows;
    int rc = kf_ensure_buf(ctx, &ctx->tmp0, &ctx->tmp0_bytes, n * 4);
    if (rc) return rc;
    void* ys[1] = {ctx->tmp0};
    rc = linear_any(ctx, 1, ys, w, x, M, KF_EPI_F32, nullptr, nullptr, 0.f);
    if (rc) return rc;
    return kf_axb_epilogue(ctx, d, (const float*)ctx->tmp0, bias, alpha, beta, w->rows, n);
}
extern "C" int kf_linear_multi(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, int M) {
    if (!ctx || !y || !w || !x) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1 && n >= 1 && n <= 3, "M, n");
    return linear_any(ctx, n, y, w, x, M, 0, nullptr, nullptr, 0.f);
}
extern "C" int kf_linear_swiglu(kf_ctx* ctx, void* y, const kf_tensor_desc* wg, const kf_tensor_desc* wu, const void* x, int M) {
    if (!ctx || !y || !wg || !wu || !x) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, M >= 1, "M");
    kf_tensor_desc w[2] = {*wg, *wu};
    void* ys[2]         = {y, y};
    return linear_any(ctx, 2, ys, w, x, M, 2, nullptr, nullptr, 0.f);
}
// RMSNorm folded into the activation staging of the matmul(s) that consume it: the normalised activations are never written to
// HBM (M <= 64).  mode: 0 = n plain outputs (n <= 3), 2 = SwiGLU(w[0] gate, w[1] up) -> y[0].
extern "C" int kf_rmsnorm_linear(kf_ctx* ctx, int n, void* const* y, const kf_tensor_desc* w, const void* x, const void* norm_w, float eps, int M,
                                 int mode) {
    if (!ctx || !y || !w || !x || !norm_w) return KF_ERR_BAD_ARG;
    KF_REQUIRE(ctx, (mode == 0 || mode == 2) && M >= 1 && n >= 1 && n <= 3, "mode / M / n");
    if (mode == 2) {
        KF_REQUIRE(ctx, n == 2, "swiglu takes gate and up");
        void* ys[2] = {y[0], y[0]};
        return linear_any(ctx, 2, ys, w, x, M, 2, nullptr, norm_w, eps);
    }
    return linear_any(ctx, n, y, w, x, M, 0, nullptr, norm_w, eps);
}
