#!/usr/bin/env python
"""Batched-decode and prefill throughput of the Qwen3 workloads of SURVEY.md 8(d) (configs 2-4) on one GPU.

    python tools/throughput_bench.py --workload qwen3-32b-q4 --batch 1,8,16,32,64 --ctx 512 --prefill 4096 --panel 2048

Batched decode: B independent sequences at position ctx (device-resident greedy loop, one CUDA-graph replay per step).
Prefill: a T-token prompt through panels of P tokens (tcgen05 linears + flash prefill attention), logits of the last token only.
One JSON line per measurement; synthetic data, random-init weights quantised at load (same generator as bench.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import koifish_b200 as kf  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def block_params(d):
    hd = 128
    qd, kd = d["n_head"] * hd, d["n_kv_head"] * hd
    return d["n_layer"] * (d["n_embd"] * (qd + 2 * kd) + qd * d["n_embd"] + 3 * d["n_embd"] * d["n_ff"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="qwen3-32b-q4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", default="1,32")
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--prefill", type=int, default=0, help="prompt length (0 = skip)")
    ap.add_argument("--panel", type=int, default=1024)
    ap.add_argument("--out", default="")
    ap.add_argument("--layers", type=int, default=0, help="debug / profiling: override the layer count (not a bench number)")
    args = ap.parse_args()
    dims_key, quantizer = WORKLOADS[args.workload]
    d = dict(kf.QWEN3_DIMS[dims_key])
    if args.layers:
        d["n_layer"] = args.layers
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = kf.Context(0, stream.cuda_stream)
    for kv in [s for s in os.environ.get("KF_SET", "").split(",") if s]:  # tuning sweeps: KF_SET=tc_min_m=9,attn_split=4
        k, v = kv.split("=")
        ctx.set_int(k, int(v))
    out = open(args.out, "a") if args.out else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def emit(rec):
        rec.update({"workload": args.workload, "data": "synthetic", "n_gpus": 1})
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")
            out.flush()

    batches = [int(b) for b in args.batch.split(",") if b]
    if batches:
        B = max(batches)
        max_seq = args.ctx + 4 * args.steps + 64
        cfg = kf.qwen3_config(d["n_layer"], d["n_embd"], d["n_ff"], d["n_head"], d["n_kv_head"], 128, 151936, quantizer, d["tie"], max_seq, B, 42, 1e6)
        model = kf.Model(ctx, cfg)
        model.init_random()
        for b in batches:
            toks = [(1000 + 37 * i) % 151936 for i in range(b)]
            model.forward(toks, [args.ctx] * b, seq_mode=1 if b > 1 else 0, want_logits=False)
            model.decode_loop(4, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launches
            profiling = bool(os.environ.get("KF_PROFILE"))  # ncu --profile-from-start off: exactly the timed steps, as single launches
            if profiling:
                model.set_graphs(False)
                torch.cuda.profiler.start()
            e0.record(stream)
            model.decode_loop(args.steps, b)
            e1.record(stream)
            torch.cuda.synchronize()
            if profiling:
                torch.cuda.profiler.stop()
                model.set_graphs(True)
            ms = e0.elapsed_time(e1) / args.steps
            emit({"mode": "decode", "batch": b, "ctx": args.ctx, "ms_per_step": round(ms, 4), "tokens_per_s": round(b * 1e3 / ms, 1),
                  "launches_per_step": (ctx.launches - l0) / args.steps, "weight_bytes": model.info.weight_bytes})
        del model
    if args.prefill:
        T, P = args.prefill, args.panel
        cfg = kf.qwen3_config(d["n_layer"], d["n_embd"], d["n_ff"], d["n_head"], d["n_kv_head"], 128, 151936, quantizer, d["tie"], T + 64, 1, 42, 1e6,
                              max_prefill=P)
        model = kf.Model(ctx, cfg)
        model.init_random()
        toks = [(1000 + 37 * i) % 151936 for i in range(T)]
        model.prefill(toks)  # eager: sizes the workspaces
        model.prefill(toks)  # captures the panel graphs
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps = 3
        profiling = bool(os.environ.get("KF_PROFILE"))
        if profiling:
            reps = 1
            model.set_graphs(False)
            torch.cuda.profiler.start()
        for _ in range(reps):
            _, nxt = model.prefill(toks)
        if profiling:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / reps
        flops = 2.0 * block_params(d) * T + 4.0 * d["n_layer"] * d["n_head"] * 128 * T * T / 2  # linears + causal attention
        emit({"mode": "prefill", "tokens": T, "panel": P, "ms": round(ms, 2), "tokens_per_s": round(T * 1e3 / ms, 1),
              "tflops": round(flops / (ms * 1e-3) / 1e12, 1), "bf16_peak_tflops_sustained": peaks.get("bf16_tflops_sustained"),
              "next_token": nxt})
    ctx.close()


if __name__ == "__main__":
    main()
