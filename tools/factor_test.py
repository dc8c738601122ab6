import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, koifish_b200 as kf, oracle_lib as ol
ctx = kf.Context(0)
N, K = 384, 4096
w = ol.fill_normal(N*K, 55, 0.02); data, gama = ol.quantize(w, N, K, 4, 128, 0)
t = kf.QTensor.from_packed(ctx, data, gama, N, K, kf.KF_T_Q4, 128, 0)
wdq = ol.dequant(data, gama, N, K, 4, 128, 0)
for M in (1, 4):
    x = ol.f32_to_bf16(np.random.default_rng(3).standard_normal((M, K)).astype(np.float32))
    ref = ol.linear_f32(wdq, x, M, N, K)
    ye = ol.bf16_to_f32(kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)).reshape(M, N)
    ctx.set_int("gemv_exact", 0)
    yf = ol.bf16_to_f32(kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)).reshape(M, N)
    ctx.set_int("gemv_exact", 1)
    s = np.sqrt((ref**2).mean())
    print("M", M, "exact rms err/rms %.2e  factor rms err/rms %.2e  max %.2e / %.2e" % (np.sqrt(((ye-ref)**2).mean())/s, np.sqrt(((yf-ref)**2).mean())/s, np.abs(ye-ref).max()/s, np.abs(yf-ref).max()/s))
