#!/usr/bin/env python
"""bench.py -- decode throughput of the quantized-inference hot path (BASELINE.json metric: decode tokens/s, Qwen3 4-bit).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload qwen3-32b-q4] [--ctx 512]

A "step" is one decode token (one pass of the hot path: 64 x [RMSNorm, fused-dequant QKV GEMV, QK-norm+RoPE+KV append, split-K
GQA attention, O GEMV+residual, RMSNorm, gate/up GEMV+SwiGLU, down GEMV+residual] + final norm + lm_head GEMV + argmax).
  value  : tokens/s with everything resident in HBM, K CUDA-graph replays timed with CUDA events on the launching stream.
  e2e    : tokens/s through the reference-facing C ABI (kf_model_forward) with HOST buffers: token id + position H2D and the
           logits + next token D2H inside the timed region, every step.
  roofline: the dominant kernel (kf_gemv_kernel, the fused unpack+dequant GEMV): algorithmic bytes per launch / its average
           launch duration (CUDA events over a graph holding exactly the step's GEMV launches), vs the measured HBM peak.
  cpu_baseline: the CPU oracle ("port") timed on this box's host cores on a bounded sample of the same workload.
N > 1: tensor parallel (one process per GPU, NCCL all-reduce per block), strong scaling of the same single-sequence decode.
--impl reference: the reference's own CPU primitives (oracle/_ref, compiled from /root/reference) assembled into the decode block.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "decode tokens/s, Qwen3 4-bit (batch 1)"
UNIT = "tokens/s"

WORKLOADS = {
    # name: (dims key, quantizer, description)
    "qwen3-32b-q4": ("32B", {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}),
    "qwen3-8b-q4": ("8B", {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}),
    "qwen3-8b-q2": ("8B", {"group_size": 128, "self_attn": {"quant_method": "yyang", "bits": 2}, "mlp": {"quant_method": "yyang", "bits": 2}}),
    "qwen3-8b-q1": ("8B", {"group_size": 128, "self_attn": {"quant_method": "yyang", "bits": 1}, "mlp": {"quant_method": "yyang", "bits": 1}}),
    "qwen3-0.6b-q4": ("0.6B", {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}),
    "qwen3-0.6b-h84": ("0.6B", {"group_size": 128, "self_attn": {"bits": 8}, "mlp": {"quant_method": "RTN", "bits": 4}}),
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


GEMV_SOURCES = ("koifish_b200/csrc/Device/gemv.cu", "koifish_b200/csrc/Device/kf_common.cuh")
TRAFFIC_FILE = "profiles/r02_traffic_decode.json"


def kernel_digest():
    """sha256 over the sources of the dominant kernel: a committed ncu traffic figure is only reported for the kernel it was taken from"""
    import hashlib
    h = hashlib.sha256()
    for rel in GEMV_SOURCES:
        h.update(open(os.path.join(ROOT, rel), "rb").read())
    return h.hexdigest()[:16]


def gemv_algorithmic_bytes(rows, cols, bits, group, M):
    """SURVEY.md 8d / BASELINE.md 3: N*K*bits/8 + (N*K/G)*4 [packed types] + 2*M*K + 2*M*N"""
    b = rows * cols * bits / 8.0 + 2.0 * M * cols + 2.0 * M * rows
    if bits in (4, 2, 1):
        b += (rows * cols / group) * 4.0
    return b


def cpu_baseline(dims, ctx_len, kind_pref="port"):
    """Bounded CPU sample of the same workload: ONE transformer block at position ctx_len-1 (oracle: packed weights dequantised at
    load, bf16 weights, fp32 accumulate, OpenMP over all host cores) plus 1/16 of the lm_head rows; extrapolated to a token."""
    import numpy as np
    import oracle_lib as ol
    t0 = time.time()
    vocab_s = 151936 // 16
    m = ol.OracleModel(n_layer=1, n_embd=dims["n_embd"], n_ff=dims["n_ff"], n_head=dims["n_head"], n_kv_head=dims["n_kv_head"], head_dim=128,
                       vocab=vocab_s, max_seq=ctx_len, rope_theta=1e6, tie_embed=1, seed=42)
    x = ol.fill_normal(dims["n_embd"], 1, 1.0)
    m.layer(0, ctx_len - 1, x)  # warm-up
    ts = []
    while len(ts) < 3 or (sum(ts) < 4.0 and len(ts) < 40):
        t = time.perf_counter()
        m.layer(0, ctx_len - 1, x)
        ts.append(time.perf_counter() - t)
    t_layer = min(ts)
    hs = []
    for _ in range(3):
        t = time.perf_counter()
        m.forward(5, 0)  # 1 layer at pos 0 + head slice
        hs.append(time.perf_counter() - t)
    t_head = max(0.0, min(hs) - t_layer) * 16
    tok_s = 1.0 / (dims["n_layer"] * t_layer + t_head)
    cores = ol.lib().kfo_num_threads()
    m.close()
    return {"value": tok_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "1 of %d blocks at pos %d (min of %d runs, %.1f ms) + 1/16 of lm_head rows (x16), extrapolated to one token; setup %.0f s"
                      % (dims["n_layer"], ctx_len - 1, len(ts), t_layer * 1e3, time.time() - t0)}


def run_reference(args, dims):
    """--impl reference: the reference's own CPU primitives (GST_float.cpp D_matvec/dotprod_fp32, rmsnorm, rope, mha_cpu) on fp32
    weights, all host threads; each step = one decode block (bounded sample), extrapolated to a token."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is meant to use every host thread it can get, and libgomp
    # reads the variable when it is loaded (below, with the oracle libraries)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle_lib as ol
    ref = None
    try:
        ol.build_oracle()
        if os.path.exists(ol.REF_SO):
            ref = C.CDLL(ol.REF_SO)
            ref.ref_decode_block_seconds.restype = C.c_double
            ref.ref_decode_block_seconds.argtypes = [C.c_int] * 7
            ref.ref_matvec_seconds.restype = C.c_double
            ref.ref_matvec_seconds.argtypes = [C.c_int] * 3
    except Exception:
        ref = None
    if ref is not None:
        # a step = ONE decode block (the bounded sample of a token): `warmup` untimed passes, then exactly `steps` timed ones, mean taken
        ref.ref_decode_block_run.restype = C.c_double
        ref.ref_decode_block_run.argtypes = [C.c_int] * 8
        t_block = ref.ref_decode_block_run(dims["n_embd"], dims["n_ff"], dims["n_head"], dims["n_kv_head"], 128, args.ctx, args.warmup, args.steps)
        t_head = ref.ref_matvec_seconds(dims["n_embd"], 151936 // 16, 2) * 16
        tok_s = 1.0 / (dims["n_layer"] * t_block + t_head)
        kind, cores = "reference", os.cpu_count()
        sample = ("each step = 1 of %d blocks at ctx %d from oracle/_ref (reference GST_float.cpp primitives, fp32 weights, OpenMP): mean of %d timed "
                  "steps after %d warm-up (%.2f ms per block); + 1/16 of the lm_head rows x16 (%.2f ms); extrapolated to one token"
                  % (dims["n_layer"], args.ctx, args.steps, args.warmup, t_block * 1e3, t_head * 1e3))
    else:
        cb = cpu_baseline(dims, args.ctx)
        tok_s, kind, cores, sample = cb["value"], "port", cb["cores"], cb["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / tok_s, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload description as our arm's line (same name, model, quantizer card, batch, context): what differs is who computes it
            "config": {"workload": "Qwen3-%s decode, batch 1, ctx %d, %s" % (WORKLOADS[args.workload][0], args.ctx, args.workload),
                       "model": "Qwen3-" + WORKLOADS[args.workload][0], "quantizer": WORKLOADS[args.workload][1], "global_batch": 1, "seq_len": args.ctx,
                       "parallelism": "host cores (the reference has no CPU inference path: its fp32 CPU primitives assembled into a block)",
                       "weights": "synthetic fp32 (the reference's CPU primitives take fp32 weights)", "layers": dims["n_layer"]},
            "cpu_baseline": {"value": tok_s, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": tok_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="qwen3-32b-q4", choices=sorted(WORKLOADS))
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--layers", type=int, default=0, help="debug: override the layer count (INVALID as a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import koifish_b200 as kf
    dims_key, quantizer = WORKLOADS[args.workload]
    dims = dict(kf.QWEN3_DIMS[dims_key])
    if args.layers:
        dims["n_layer"] = args.layers
    if args.impl == "reference":
        return run_reference(args, dims)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()          # a real (non-default) stream: torch events and our kernels share it
    torch.cuda.set_stream(stream)
    ctx = kf.Context(local, stream.cuda_stream)
    for knob in ("tp_fused", "attn_split", "gemv_splitk", "pdl", "gemv_exact", "tc_min_m", "attn_warps", "gemv_cluster", "gqa_min_ctx", "gemv_tma", "gemv_tma_occ",
                 "gemv_tma_smem_kb", "deq_fma"):  # tuning experiments only, e.g. KF_ATTN_SPLIT=16
        if os.environ.get("KF_" + knob.upper()):
            ctx.set_int(knob, int(os.environ["KF_" + knob.upper()]))
    gemv_exact = ctx.get_int("gemv_exact")
    if world > 1:
        ctx.init_tensor_parallel(rank, world, p2p=os.environ.get("KF_P2P", "1") != "0")  # KF_P2P=0: NCCL exchange only, for comparison

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    max_seq = max(1024, args.ctx + args.steps * 2 + args.warmup * 2 + 64)
    cfg = kf.qwen3_config(dims["n_layer"], dims["n_embd"], dims["n_ff"], dims["n_head"], dims["n_kv_head"], 128, 151936, quantizer, dims["tie"],
                          max_seq, 1, 42, 1e6)
    t_setup = time.time()
    model = kf.Model(ctx, cfg, rank, world)
    model.init_random()
    info = model.info
    # ---- context: fill the KV cache through the real path (prefill panels of 64 tokens) ------------------------------------
    toks = [(1000 + 37 * i) % 151936 for i in range(args.ctx)]
    for p0 in range(0, args.ctx - 1, 64):
        p1 = min(args.ctx - 1, p0 + 64)
        model.forward(toks[p0:p1], list(range(p0, p1)), seq_mode=0, want_logits=False)
    _, nxt = model.forward([toks[args.ctx - 1]], [args.ctx - 1], want_logits=True, want_next=True)
    t_setup = time.time() - t_setup

    # ---- (1) device-resident decode: value --------------------------------------------------------------------------------------
    pos = args.ctx
    model.forward([int(nxt[0])], [pos], want_logits=False)      # stage (token, pos) for the loop
    model.decode_loop(args.warmup, 1)                           # W untimed steps (first eager, then the captured graph)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    profiling = bool(os.environ.get("KF_PROFILE"))  # ncu --profile-from-start off: capture exactly the timed decode steps
    if profiling:
        model.set_graphs(False)                      # individual launches instead of one graph node
        torch.cuda.profiler.start()
    e0.record(stream)
    model.decode_loop(args.steps, 1)
    e1.record(stream)
    barrier()
    if profiling:
        torch.cuda.profiler.stop()
        model.set_graphs(True)
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    launches = ctx.launches - l0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    _, pos_now = model.read_state(1)
    pos = int(pos_now[0])

    # ---- (2) end to end through the C ABI with host buffers ------------------------------------------------------------------------
    tok = int(model.read_state(1)[0][0])
    for _ in range(args.warmup):
        lg, nx = model.forward([tok], [pos], want_logits=True, want_next=True)
        tok, pos = int(nx[0]), pos + 1
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host = time.perf_counter()
    e2.record(stream)
    for _ in range(args.steps):
        lg, nx = model.forward([tok], [pos], want_logits=True, want_next=True)
        tok, pos = int(nx[0]), pos + 1
    e3.record(stream)
    barrier()
    t_host = time.perf_counter() - t_host
    ms2 = torch.tensor([max(e2.elapsed_time(e3), t_host * 1e3)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_tok_s = args.steps / (float(ms2.item()) / 1e3)
    clocks = sampler.stop()

    # ---- (3) roofline of the dominant kernel: a graph with exactly the step's GEMV launches -----------------------------------------
    E, F, hd = info.n_embd, info.n_ff // world, 128
    QD, KD = (info.n_head // world) * hd, (info.n_head_kv // world) * hd
    xbuf = kf.fill_normal(ctx, max(E, F, QD) * 1, 99, 1.0)
    ybufs = [ctx.empty(max(E, F, QD, 151936) * 4) for _ in range(3)]
    plan, alg_bytes = [], 0.0
    bits_of = lambda d: kf.TYPE_BITS[d.type]  # noqa: E731
    for l in range(info.n_layers):
        p = "model.layers.%d." % l
        dq, dk, dv = (model.tensor_desc(p + "self_attn.%s_proj.weight" % n) for n in "qkv")
        do = model.tensor_desc(p + "self_attn.o_proj.weight")
        dg, du, dd = (model.tensor_desc(p + "mlp.%s_proj.weight" % n) for n in ("gate", "up", "down"))
        plan.append(("multi", (dq, dk, dv)))
        plan.append(("lin", do))
        plan.append(("swiglu", (dg, du)))
        plan.append(("lin", dd))
        for d in (dq, dk, dv, do, dg, du, dd):
            alg_bytes += gemv_algorithmic_bytes(d.rows, d.cols, bits_of(d), d.group, 1)
    # the bf16 lm_head goes through the TMA-fed tcgen05 kernel (HBM-roofline already, profiles/r01_tc_crossover.txt): it is counted in
    # the step's bytes but not in the roofline of the dominant kernel, the block GEMVs
    dh = model.tensor_desc("lm_head.weight" if not info.tie_word_embeddings else "model.embed_tokens.weight")
    head_rows = dh.rows // world
    head_bytes = gemv_algorithmic_bytes(head_rows, dh.cols, bits_of(dh), dh.group, 1)
    n_gemv = len(plan)

    def gemv_pass():
        lib, h = ctx.lib, ctx.h
        for kind, d in plan:
            if kind == "multi":
                descs = (kf.TensorDesc * 3)(*d)
                ys = (C.c_void_p * 3)(ybufs[0].ptr, ybufs[1].ptr, ybufs[2].ptr)
                ctx.check(lib.kf_linear_multi(h, 3, ys, descs, xbuf.ptr, 1), "kf_linear_multi")
            elif kind == "swiglu":
                ctx.check(lib.kf_linear_swiglu(h, ybufs[0].ptr, C.byref(d[0]), C.byref(d[1]), xbuf.ptr, 1), "kf_linear_swiglu")
            else:
                ctx.check(lib.kf_linear(h, ybufs[0].ptr, C.byref(d), xbuf.ptr, 1, 0, None), "kf_linear")

    gemv_pass()  # eager: sizes the workspaces
    g = C.c_void_p()
    ctx.check(ctx.lib.kf_graph_begin(ctx.h), "kf_graph_begin")
    gemv_pass()
    ctx.check(ctx.lib.kf_graph_end(ctx.h, C.byref(g)), "kf_graph_end")
    for _ in range(3):
        ctx.check(ctx.lib.kf_graph_launch(ctx.h, g), "kf_graph_launch")
    torch.cuda.synchronize()
    reps = max(4, min(16, args.steps // 4))
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record(stream)
    for _ in range(reps):
        ctx.check(ctx.lib.kf_graph_launch(ctx.h, g), "kf_graph_launch")
    e5.record(stream)
    torch.cuda.synchronize()
    gemv_ms_pass = e4.elapsed_time(e5) / reps
    ctx.lib.kf_graph_destroy(g)
    peak, peak_src = read_peaks()
    per_launch_bytes = alg_bytes / n_gemv
    per_launch_s = gemv_ms_pass / 1e3 / n_gemv
    achieved = per_launch_bytes / per_launch_s / 1e9
    kv_bytes_tok = 2.0 * info.n_layers * (args.ctx + args.steps // 2) * KD * 2
    step_alg_bytes = alg_bytes + head_bytes + kv_bytes_tok
    # measured DRAM bytes per launch of the same kernel in the same decode step (one ncu pass, tools/gpu_traffic.sh), if committed
    traffic, traffic_src = None, None
    if world == 1 and args.workload == "qwen3-32b-q4" and not args.layers:
        try:
            tf = json.load(open(os.path.join(ROOT, TRAFFIC_FILE)))
            if tf.get("kernel_digest") != kernel_digest() or tf.get("gemv_exact") != gemv_exact:
                traffic_src = "%s is stale (taken from kernel digest %s / gemv_exact %s, this build is %s / %s): not reported" % (
                    TRAFFIC_FILE, tf.get("kernel_digest"), tf.get("gemv_exact"), kernel_digest(), gemv_exact)
            else:
                t = tf["kernels"]["kf_gemv_kernel"]
                traffic = t["avg_dram_read_bytes"] + t["avg_dram_write_bytes"]
                traffic_src = "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, avg of %d launches, kernel digest %s)" % (
                    TRAFFIC_FILE, t["launches"], tf["kernel_digest"])
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "kernel": "kf_gemv_kernel (fused unpack+dequant GEMV, M=1, %s)" % ("in-kernel dequant to the reference's bf16 weights" if gemv_exact else "fp16-code tensor-core arithmetic, group affine on fp32 sums"), "launches_per_step": n_gemv,
                "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_us": per_launch_s * 1e6, "gemv_share_of_step": gemv_ms_pass / ms_step,
                "peak_source": peak_src, "frac_of_8TBps": achieved / 8000.0,
                "step_gbps_all_kernels": step_alg_bytes / (ms_step / 1e3) / 1e9, "step_frac_of_peak": step_alg_bytes / (ms_step / 1e3) / 1e9 / peak}

    if rank == 0:
        cb = None if args.no_cpu_baseline or world > 1 else cpu_baseline(dims, args.ctx)
        line = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Qwen3-%s decode, batch 1, ctx %d, %s" % (dims_key, args.ctx, args.workload), "model": "Qwen3-" + dims_key,
                       "quantizer": quantizer, "global_batch": 1, "seq_len": args.ctx, "parallelism": "tp%d" % world,
                       "weights": "random-init N(0,0.02^2)-like, quantised at load on the GPU", "lm_head": "bf16",
                       "gemv_exact": gemv_exact,
                       "l2": "weights per token (%.1f GB) exceed the 126 MB L2; no reuse between steps" % ((alg_bytes + head_bytes) / 1e9),
                       "layers": info.n_layers, "setup_s": round(t_setup, 1)},
            "e2e": {"value": e2e_tok_s, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 151936 * 2 + 4},
            "gpu_launches": int(launches), "launches_per_step": launches / max(1, args.steps),
            "roofline": roofline, "clocks": clocks,
            "weight_bytes_per_rank": int(info.weight_bytes),
        }
        if cb:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    model.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
