#!/usr/bin/env python
"""Config 1 in full (Qwen3-0.6B, 28 layers, vocab 151936, 4-bit RTN blocks): 16-token prompt + 128 greedy tokens chosen by the CPU oracle and
teacher-forced into the GPU path; prints, per arithmetic mode, the distribution of the logits error and the top-1 agreement.
    python tools/greedy_gate.py [--steps 128] [--theta 1e6]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import koifish_b200 as kf  # noqa: E402
import oracle_lib as ol  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--theta", type=float, default=1e6)
    ap.add_argument("--layers", type=int, default=28)
    args = ap.parse_args()
    V = 151936
    quantizer = {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}
    oracle = ol.OracleModel(n_layer=args.layers, n_embd=1024, n_ff=3072, n_head=16, n_kv_head=8, head_dim=128, vocab=V, max_seq=512,
                            rope_theta=args.theta, tie_embed=1, seed=42, norm_sigma=0.1)
    toks = [(1000 + 37 * i) % V for i in range(16)]
    wants = []
    pos = 0
    while pos < len(toks):
        w = ol.bf16_to_f32(oracle.forward(toks[pos], pos))
        wants.append(w)
        if pos >= 15 and len(toks) < 16 + args.steps:
            toks.append(int(np.argmax(w)))
        pos += 1
    ctx = kf.Context(0)
    for exact in (1, 0):
        ctx.set_int("gemv_exact", exact)
        cfg = kf.qwen3_config(args.layers, 1024, 3072, 16, 8, 128, V, quantizer, True, 512, 1, 42, args.theta, norm_sigma=0.1)
        model = kf.Model(ctx, cfg)
        model.init_random()
        errs, rmss, same, gaps = [], [], [], []
        for pos, tok in enumerate(toks):
            lg, _ = model.forward([tok], [pos])
            g, w = ol.bf16_to_f32(lg[0]), wants[pos]
            errs.append(float(np.abs(g - w).max() / np.abs(w).max()))
            rmss.append(float(np.sqrt(np.mean((g - w) ** 2)) / np.sqrt(np.mean(w ** 2))))
            top2 = np.sort(w)[-2:]
            gaps.append(float((top2[1] - top2[0]) / np.abs(w).max()))
            same.append(int(np.argmax(g) == np.argmax(w)))
        errs, rmss, same, gaps = map(np.array, (errs, rmss, same, gaps))
        print("gemv_exact=%d theta=%g layers=%d: %d positions | max err / max|logit|: worst %.3e median %.3e | rms err / rms logit: worst %.3e median %.3e | "
              "top-1 equal %d/%d | wrong top-1 only where the oracle's top-2 gap is <= %.3e of the largest logit"
              % (exact, args.theta, args.layers, len(toks), errs.max(), np.median(errs), rmss.max(), np.median(rmss), same.sum(), len(toks),
                 gaps[same == 0].max() if (same == 0).any() else 0.0), flush=True)
        model.close()
    ctx.close()


if __name__ == "__main__":
    main()
