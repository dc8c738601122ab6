"""CPU tests of the tensor-parallel host logic with a real 2-process gloo group (no GPU): every rank asks the C ABI for its shard
windows; together the windows must tile each tensor exactly once, quant groups must never straddle shards, and the synthetic
weights of a shard must equal the slice of the full tensor (checked with the oracle's generator)."""
import ctypes as C
import json
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NAMES = ["model.embed_tokens.weight", "model.norm.weight", "lm_head.weight", "model.layers.3.input_layernorm.weight",
         "model.layers.3.self_attn.q_proj.weight", "model.layers.3.self_attn.k_proj.weight", "model.layers.3.self_attn.v_proj.weight",
         "model.layers.3.self_attn.o_proj.weight", "model.layers.3.self_attn.q_norm.weight", "model.layers.3.mlp.gate_proj.weight",
         "model.layers.3.mlp.up_proj.weight", "model.layers.3.mlp.down_proj.weight"]


def _shard(lib, cfg_text, name, rank, world):
    out = (C.c_int * 6)()
    err = C.c_void_p()
    st = lib.kf_config_shard_of(cfg_text, name.encode(), rank, world, out, C.byref(err))
    if err.value:
        lib.kf_string_free(err)
    return st, list(out)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import koifish_b200 as kf
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = kf.load()
        cfg = json.dumps(kf.qwen3_config(**kf.QWEN3_DIMS["32B"])).encode()
        mine = []
        for n in NAMES:
            st, s = _shard(lib, cfg, n, rank, world)
            assert st == 0, n
            mine.append(s)
        t = torch.tensor(mine, dtype=torch.int64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if rank == 0:
            q.put([g.tolist() for g in gathered])
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_shard_plan_tiles_every_tensor_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    plans = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, name in enumerate(NAMES):
        rows_g, cols_g = plans[0][i][0], plans[0][i][1]
        cover = np.zeros((min(rows_g, 4096), min(cols_g, 4096)), dtype=np.int32)  # a corner is enough to catch overlaps
        sharded = any(plans[r][i][2:4] != [rows_g, cols_g] for r in range(world))
        for r in range(world):
            rg, cg, rl, cl, r0, c0 = plans[r][i]
            assert (rg, cg) == (rows_g, cols_g)
            if sharded:
                assert cl == cg or (cl % 128 == 0 and c0 % 128 == 0), name  # 128-wide quant groups never straddle shards
                cover[r0:min(r0 + rl, cover.shape[0]), c0:min(c0 + cl, cover.shape[1])] += 1
        if sharded:
            assert rl * cl * world == rows_g * cols_g
            full = np.zeros_like(cover)
            for r in range(world):
                _, _, rl, cl, r0, c0 = plans[r][i]
            assert cover.max() <= 1
        else:
            assert all(plans[r][i][2:] == [rows_g, cols_g, 0, 0] for r in range(world))


def test_shard_plan_rules_and_errors():
    import koifish_b200 as kf
    lib = kf.load()
    cfg = json.dumps(kf.qwen3_config(**kf.QWEN3_DIMS["32B"])).encode()
    # 32B at TP = 8: 8 Q heads + 1 KV head per rank, FFN 3200, O/down split along K in multiples of 128
    assert _shard(lib, cfg, "model.layers.0.self_attn.q_proj.weight", 3, 8) == (0, [8192, 5120, 1024, 5120, 3072, 0])
    assert _shard(lib, cfg, "model.layers.0.self_attn.k_proj.weight", 7, 8) == (0, [1024, 5120, 128, 5120, 896, 0])
    assert _shard(lib, cfg, "model.layers.0.self_attn.o_proj.weight", 1, 8) == (0, [5120, 8192, 5120, 1024, 0, 1024])
    assert _shard(lib, cfg, "model.layers.0.mlp.down_proj.weight", 2, 4) == (0, [5120, 25600, 5120, 6400, 0, 12800])
    assert _shard(lib, cfg, "model.layers.0.mlp.gate_proj.weight", 0, 2)[1][2:] == [12800, 5120, 0, 0]
    assert _shard(lib, cfg, "model.norm.weight", 1, 2) == (0, [1, 5120, 1, 5120, 0, 0])
    # indivisible degree (8 KV heads) and unknown names are errors, not silent fallbacks
    assert _shard(lib, cfg, "model.layers.0.self_attn.q_proj.weight", 0, 3)[0] != 0
    assert _shard(lib, cfg, "model.layers.0.self_attn.q_proj.weight", 0, 16)[0] != 0
    assert _shard(lib, cfg, "model.layers.0.bogus.weight", 0, 2)[0] != 0


def test_shard_of_synthetic_weights_equals_slice_of_full_tensor():
    # the generator index of element (r, c) of a shard is its index in the FULL tensor (kf_fill_normal_2d); restated with the oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    rows_g, cols_g, world = 64, 512, 4
    full = ol.fill_normal(rows_g * cols_g, 99, 0.02).reshape(rows_g, cols_g)
    for rank in range(world):
        cl = cols_g // world
        shard = full[:, rank * cl:(rank + 1) * cl]
        # quantising the shard == slicing the quantised full tensor (groups of 128 along K stay intact)
        d_full, g_full = ol.quantize(full.reshape(-1), rows_g, cols_g, 4, 128, ol.RTN_ASYM)
        d_sh, g_sh = ol.quantize(np.ascontiguousarray(shard).reshape(-1), rows_g, cl, 4, 128, ol.RTN_ASYM)
        dq_full = ol.dequant(d_full, g_full, rows_g, cols_g, 4, 128, 0)
        dq_sh = ol.dequant(d_sh, g_sh, rows_g, cl, 4, 128, 0)
        assert np.array_equal(dq_sh, dq_full[:, rank * cl:(rank + 1) * cl])
