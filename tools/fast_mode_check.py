#!/usr/bin/env python
"""Accuracy of the decode GEMV's arithmetic modes against the CPU oracle (fp32 accumulation over the reference's dequantised weights):
exact (bit-exact bf16 dequant inside the kernel, gemv_exact=1) vs fast (fp16 codes + affine map on the group sums, gemv_exact=0).
Covers residual / SwiGLU / fused-norm / multi-weight paths and wide dynamic ranges of the activations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import koifish_b200 as kf  # noqa: E402
import oracle_lib as ol  # noqa: E402

ctx = kf.Context(0)
if os.environ.get("KF_GEMV_TMA"): ctx.set_int("gemv_tma", int(os.environ["KF_GEMV_TMA"]))
rng = np.random.default_rng(3)
worst = {0: 0.0, 1: 0.0}
for (N, K) in ((384, 4096), (1024, 5120), (5120, 8192), (640, 25600)):
    w = ol.fill_normal(N * K, 55 + N, 0.02)
    for mode, qb in ((ol.RTN_ASYM, 0), (ol.RTN_SYM, 8)):
        data, gama = ol.quantize(w, N, K, 4, 128, mode)
        t = kf.QTensor.from_packed(ctx, data, gama, N, K, kf.KF_T_Q4, 128, qb)
        wdq = ol.dequant(data, gama, N, K, 4, 128, qb)
        for M in (1, 3, 8):
            for xscale in (1.0, 300.0, 1e-4):
                x = rng.standard_normal((M, K)).astype(np.float32) * xscale
                x[:, ::97] *= 50.0  # outliers
                x[:, 5::31] *= 1e-5  # tiny values next to large ones
                xb = ol.f32_to_bf16(x)
                ref = ol.linear_f32(wdq, xb, M, N, K)
                s = np.sqrt((ref ** 2).mean())
                out = []
                for exact in (1, 0):
                    ctx.set_int("gemv_exact", exact)
                    y = ol.bf16_to_f32(kf.linear(ctx, t, ctx.array(xb), M).numpy(np.uint16)).reshape(M, N)
                    err = np.abs(y - ref).max() / s
                    worst[exact] = max(worst[exact], err)
                    out.append(err)
                print("N %5d K %5d qbias %d M %d xscale %-7g  max|err|/rms: exact %.2e  fast %.2e" % (N, K, qb, M, xscale, out[0], out[1]), flush=True)
ctx.set_int("gemv_exact", 1)
print("WORST exact %.3e fast %.3e" % (worst[1], worst[0]))
assert worst[0] < 2e-2 and worst[1] < 2e-2
print("OK")
