#!/usr/bin/env python
"""Tensor-parallel parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py
Every rank builds its shard of the same synthetic model (sharded at pack time: the packed shard equals the slice of the full
packed tensor), decodes the same tokens with an NCCL all-reduce after O and down, and rank 0 compares the logits with the
unsharded (TP = 1) model built on its own GPU.  Exit code 0 = parity."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import koifish_b200 as kf  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = kf.Context(local)
    ctx.init_tensor_parallel(rank, world)  # NCCL communicator + peer-memory exchange buffers
    if rank == 0:
        print("tp_check: peer-memory exchange %s" % ("active" if ctx.lib.kf_p2p_ready(ctx.h) == 1 else "NOT available, NCCL path"), flush=True)

    quantizer = {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}
    # 8 KV heads so that TP up to 8 divides; small enough to build twice on rank 0
    cfg = kf.qwen3_config(4, 1024, 4096, 16, 8, 128, 16384, quantizer, False, 128, 2, 42, 1e6, norm_sigma=0.1)
    model = kf.Model(ctx, cfg, rank, world)
    model.init_random()
    ref = None
    if rank == 0:
        ctx1 = kf.Context(local)
        ref = kf.Model(ctx1, cfg, 0, 1)
        ref.init_random()
    toks = [(1000 + 37 * i) % 16384 for i in range(12)]
    worst, ok = 0.0, True
    fused = ctx.lib.kf_exchange_fused_ready(ctx.h, 1, 1024) == 1
    if rank == 0:
        print("tp_check: exchange fused into the O / down epilogues: %s" % ("yes" if fused else "NO"), flush=True)
    # (a) the same tokens with the stand-alone exchange kernel (knob tp_fused = 0): the fused path adds the partials in the same rank order
    #     and rounds at the same points, so the logits must be BIT-identical -- eager, captured graph and replays alike
    unfused = []
    ctx.set_int("tp_fused", 0)
    m0 = kf.Model(ctx, cfg, rank, world)
    m0.init_random()
    for pos, tok in enumerate(toks):
        lg, _ = m0.forward([tok], [pos], want_logits=True)
        unfused.append(lg[0].copy())
    lg2u, _ = m0.forward([5, 9], [12, 12], seq_mode=1)
    m0.close()
    ctx.set_int("tp_fused", 1)
    same_bits = True
    for pos, tok in enumerate(toks):
        lg, nx = model.forward([tok], [pos], want_logits=True, want_next=True)
        same_bits = same_bits and bool(np.array_equal(lg[0], unfused[pos]))
        if rank == 0:
            lr, nr = ref.forward([tok], [pos], want_logits=True, want_next=True)
            a = (lg[0].astype(np.uint32) << 16).view(np.float32)
            b = (lr[0].astype(np.uint32) << 16).view(np.float32)
            err = float(np.abs(a - b).max() / np.abs(b).max())
            worst = max(worst, err)
    # batched decode (2 independent sequences) and a prefill panel through the sharded path
    lg2, _ = model.forward([5, 9], [12, 12], seq_mode=1)
    same_bits = same_bits and bool(np.array_equal(lg2, lg2u))
    if fused and not same_bits:
        print("tp_check: rank %d: fused exchange differs from the stand-alone exchange kernel" % rank, flush=True)
        ok = False
    lgp, _ = model.forward(toks[:8], list(range(20, 28)), seq_mode=0)
    if rank == 0:
        l2, _ = ref.forward([5, 9], [12, 12], seq_mode=1)
        lp, _ = ref.forward(toks[:8], list(range(20, 28)), seq_mode=0)
        for x, y in ((lg2, l2), (lgp, lp)):
            a = (x.astype(np.uint32) << 16).view(np.float32)
            b = (y.astype(np.uint32) << 16).view(np.float32)
            worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
        ok = ok and worst <= 1e-2
        print("tp_check: world %d  worst logits rel err vs TP=1: %.3e  -> %s" % (world, worst, "OK" if ok else "FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    model.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
