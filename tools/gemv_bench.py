#!/usr/bin/env python
"""Microbench sweep of the fused dequant GEMV / skinny GEMM (BASELINE.json config 5):
    y[M,N] = x[M,K] . deq(W[N,K])^T,  K,N in {4096, 5120, 25600}, bits {16, 8, 4, 2, 1}, M in {1, 16, ...}
Reports achieved algorithmic GB/s against the measured HBM peak and 8 TB/s.  Weights rotate through buffers larger than the
126 MB L2 so every launch streams from HBM.  One JSON line per case (also appended to --out).

    python tools/gemv_bench.py [--quick] [--splitk 0,1,2,4] [--out gpurun_out/gemv_sweep.jsonl]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import koifish_b200 as kf  # noqa: E402

TYPES = {"bf16": (kf.KF_T_BF16, 0), "f8": (kf.KF_T_F8E5M2, 0), "q4": (kf.KF_T_Q4, kf.KF_Q_RTN_ASYM), "q2t": (kf.KF_T_SIGN, kf.KF_Q_YYANG),
         "q1": (kf.KF_T_BINARY, kf.KF_Q_YYANG)}


def alg_bytes(N, K, bits, M, group=128):
    b = N * K * bits / 8.0 + 2.0 * M * K + 2.0 * M * N
    if bits in (4, 2, 1):
        b += (N * K / group) * 4.0
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--splitk", default="0")
    ap.add_argument("--types", default="q4,q2t,q1,f8,bf16")
    ap.add_argument("--ms", default="1,16")
    ap.add_argument("--shapes", default="")
    ap.add_argument("--variants", default="0", help="gemv_variant values: 0 auto, 1 = 16 rows/warp, 2 = 32 rows/warp")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="")
    ap.add_argument("--tc", default="-1", help="tc_min_m values: token count from which the tcgen05 GEMM is used (-1 = auto per type, 0 = never, 1 = always)")
    ap.add_argument("--set", default="", help="extra context knobs, e.g. gemv_cluster=0,attn_split=4")
    ap.add_argument("--exact", type=int, default=1, help="gemv_exact knob: 1 = reference-exact in-kernel dequant, 0 = factored scale/zero")
    args = ap.parse_args()
    stream = torch.cuda.Stream()  # a real (non-default) stream shared by torch events and our kernels
    torch.cuda.set_stream(stream)
    ctx = kf.Context(0, stream.cuda_stream)
    ctx.set_int("gemv_exact", args.exact)
    for kv in [s for s in args.set.split(",") if s]:
        k, v = kv.split("=")
        ctx.set_int(k, int(v))
    peak = 6452.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if args.shapes:
        shapes = [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")]
    elif args.quick:
        shapes = [(51200, 5120), (5120, 25600), (4096, 4096)]
    else:
        shapes = [(n, k) for n in (4096, 5120, 25600) for k in (4096, 5120, 25600)] + [(51200, 5120), (10240, 5120), (5120, 8192), (151936, 5120)]
    out = open(args.out, "a") if args.out else None
    for tname in args.types.split(","):
        tp, mode = TYPES[tname]
        bits = kf.TYPE_BITS[tp]
        for (N, K) in shapes:
            wbytes = N * K * bits / 8
            nbuf = max(2, min(24, int(400e6 // wbytes) + 1))
            src = kf.fill_normal(ctx, N * K, 1234, 0.02)
            ws = [kf.quantize(ctx, src, N, K, tp, 128, mode) for _ in range(nbuf)]
            del src
            for M in [int(m) for m in args.ms.split(",")]:
                x = kf.fill_normal(ctx, M * K, 7, 1.0)
                y = ctx.empty(M * N * 2)
                for sk, variant, tc in [(int(s), int(v), int(t)) for s in args.splitk.split(",") for v in args.variants.split(",") for t in args.tc.split(",")]:
                    ctx.set_int("tc_min_m", tc)
                    ctx.set_int("gemv_splitk", sk)
                    ctx.set_int("gemv_variant", variant)
                    descs = [w.desc() for w in ws]
                    for i in range(3):
                        ctx.check(ctx.lib.kf_linear(ctx.h, y.ptr, C.byref(descs[i % nbuf]), x.ptr, M, 0, None), "kf_linear")
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    # the launches go into ONE CUDA graph: a ~10 us kernel issued from Python is otherwise timed at the host's launch rate
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream):
                        for i in range(args.iters):
                            ctx.check(ctx.lib.kf_linear(ctx.h, y.ptr, C.byref(descs[i % nbuf]), x.ptr, M, 0, None), "kf_linear")
                    g.replay()
                    torch.cuda.synchronize()
                    e0.record(stream)
                    g.replay()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / args.iters
                    ab = alg_bytes(N, K, bits, M)
                    gbs = ab / (us * 1e-6) / 1e9
                    rec = {"type": tname, "N": N, "K": K, "M": M, "tc": tc, "splitk": sk, "S": ctx.get_int("gemv_last_s"), "variant": variant, "us": round(us, 2), "alg_MB": round(ab / 1e6, 2), "GBps": round(gbs, 1),
                           "frac_measured": round(gbs / peak, 3), "frac_8TBps": round(gbs / 8000.0, 3), "tflops": round(2.0 * M * N * K / (us * 1e-6) / 1e12, 2),
                           "nbuf": nbuf}
                    print(json.dumps(rec), flush=True)
                    if out:
                        out.write(json.dumps(rec) + "\n")
                        out.flush()
                ctx.set_int("gemv_splitk", 0)
                ctx.set_int("gemv_variant", 0)
            del ws
    ctx.close()


if __name__ == "__main__":
    main()
