"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares; host-side config / quant-card logic;
the product fails loudly without a device (no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import koifish_b200 as kf
from koifish_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("kf_device.h", "kf_model.h", "kf_tokenizer.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(kf_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    lib = kf.load()
    declared = _declared_symbols()
    assert len(declared) > 50
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    # and the Python binding covers exactly the declared surface
    assert set(L.SIGNATURES) == declared


def test_headers_compile_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "kf_device.h"\n#include "kf_model.h"\n#include "kf_tokenizer.h"\nint main(void){ kf_tensor_desc d; (void)d; return KF_OK; }\n')
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_status_strings_and_no_cpu_fallback():
    lib = kf.load()
    assert lib.kf_status_string(0) == b"KF_OK"
    assert b"no CPU fallback" in lib.kf_status_string(kf.KF_ERR_NO_DEVICE)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(kf.KoifishError) as e:
            kf.Context(0)
        assert e.value.status == kf.KF_ERR_NO_DEVICE


def _quant_of(cfg, name):
    lib = kf.load()
    t, g, m, qb, err = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
    st = lib.kf_config_quant_of(json.dumps(cfg).encode(), name.encode(), C.byref(t), C.byref(g), C.byref(m), C.byref(qb), C.byref(err))
    msg = C.cast(err, C.c_char_p).value.decode() if err.value else ""
    if err.value:
        lib.kf_string_free(err)
    return st, t.value, g.value, m.value, qb.value, msg


REF_Q4_QUANTIZER = {  # the quantizer block of the reference's cases/qwen3/qwen3_596M_q4.json
    "#MIQ": ["self_attn"], "train_target": "gama", "group_size": 128,
    "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4},
    "# embed_tokens": {"bits": 4}, "#MINI": "off"}


def test_quant_card_selection_matches_reference_rules():
    cfg = kf.qwen3_config(6, 1024, 3072, 16, 8, quantizer=REF_Q4_QUANTIZER, tie=True)
    st, t, g, m, qb, _ = _quant_of(cfg, "model.layers.3.self_attn.q_proj.weight")
    assert (st, t, g, m, qb) == (0, kf.KF_T_Q4, 128, kf.KF_Q_RTN_ASYM, 0)
    st, t, g, m, qb, _ = _quant_of(cfg, "model.layers.0.mlp.down_proj.weight")
    assert (st, t, g) == (0, kf.KF_T_Q4, 128)
    # '#'-prefixed keys are comments: embed_tokens stays bf16
    st, t, *_ = _quant_of(cfg, "model.embed_tokens.weight")
    assert (st, t) == (0, kf.KF_T_BF16)
    # hybrid 8/4-bit, ternary and binary cards
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"bits": 8}, "mlp": {"quant_method": "RTN", "bits": 4, "group_size": 256}})
    assert _quant_of(cfg, "model.layers.1.self_attn.k_proj.weight")[1] == kf.KF_T_F8E5M2
    assert _quant_of(cfg, "model.layers.1.mlp.up_proj.weight")[1:3] == (kf.KF_T_Q4, 256)
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"quant_method": "yyang", "bits": 2}, "mlp": {"quant_method": "yyang", "bits": 1}})
    st, t, g, m, qb, _ = _quant_of(cfg, "model.layers.1.self_attn.q_proj.weight")
    assert (t, m, qb) == (kf.KF_T_SIGN, kf.KF_Q_YYANG, 1)
    st, t, g, m, qb, _ = _quant_of(cfg, "model.layers.1.mlp.gate_proj.weight")
    assert (t, m, qb) == (kf.KF_T_BINARY, kf.KF_Q_YYANG, 0)


def test_bits_without_method_selects_normalfloat4():
    # {"bits": 4} with no quant_method: QUANT_MODE::RTNf (GeQuant.cpp:1270-1280) -> typNUMBER::Q4 carrying NormalFloat4 codes + per-row codebooks
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"bits": 4}, "mlp": {"bits": 8}})
    st, t, *_ = _quant_of(cfg, "model.layers.0.self_attn.q_proj.weight")
    assert (st, t) == (0, kf.KF_T_NF4)
    assert _quant_of(cfg, "model.layers.0.mlp.up_proj.weight")[1] == kf.KF_T_F8E5M2
    # bits outside {1, 2, 4, 8} fall back to 4 (Init4Neuron, GeQuant.cpp:1253-1256: `default_bits = 4; assert(0)` -- a release build carries on)
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"bits": 3}})
    assert _quant_of(cfg, "model.layers.0.self_attn.q_proj.weight")[:2] == (0, kf.KF_T_NF4)
    # bits 2 without a method is not a reference mode (RT_NormalF asserts 4 or 3 bits)
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"bits": 2}})
    st, *_, msg = _quant_of(cfg, "model.layers.0.self_attn.q_proj.weight")
    assert st == kf.KF_ERR_UNSUPPORTED and msg


def test_out_of_scope_quant_methods_fail_loudly():
    for q in ({"self_attn": {"quant_method": "awq", "bits": 8}},  # the vendor AWQ layout is 4-bit / group 128 only
              {"self_attn": {"quant_method": "awq", "bits": 4, "group_size": 64}},
              {"self_attn": {"quant_method": "bitnet"}},
              {"self_attn": {"quant_method": "RTN", "bits": 1}}):
        cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer=q)
        st, *_, msg = _quant_of(cfg, "model.layers.0.self_attn.q_proj.weight")
        assert st == kf.KF_ERR_UNSUPPORTED and msg


def test_vendor_awq_card_and_hf_quantization_config():
    # {"quant_method": "awq"} selects typNUMBER::Q4 under QUANT_MODE::AWQ (Init4Neuron, GeQuant.cpp:1272-1273) = the library's KF_T_AWQ4
    cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer={"self_attn": {"quant_method": "awq", "bits": 4, "zero_point": True}})
    st, t, g, m, qb, msg = _quant_of(cfg, "model.layers.0.self_attn.q_proj.weight")
    assert (st, t, g, qb) == (0, kf.KF_T_AWQ4, 128, 0), msg
    assert _quant_of(cfg, "model.layers.0.mlp.up_proj.weight")[1] == kf.KF_T_BF16
    # an HF config.json of a vendor checkpoint (Qwen3-32B-AWQ): QUANT_CARD::Vendor2JSONx (CLI_params.cpp:240-262) spreads the vendor's block
    # over every self_attn / mlp linear; embeddings, norms and the head stay bf16
    hf = {"hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 2, "num_attention_heads": 16, "num_key_value_heads": 8,
          "head_dim": 128, "vocab_size": 151936, "rope_theta": 1e6, "model_type": "qwen3",
          "quantization_config": {"bits": 4, "group_size": 128, "modules_to_not_convert": None, "quant_method": "awq", "version": "gemm",
                                  "zero_point": True}}
    for name in ("model.layers.1.self_attn.o_proj.weight", "model.layers.0.mlp.down_proj.weight"):
        st, t, g, *_, msg = _quant_of(hf, name)
        assert (st, t, g) == (0, kf.KF_T_AWQ4, 128), msg
    for name in ("model.embed_tokens.weight", "lm_head.weight", "model.norm.weight"):
        assert _quant_of(hf, name)[:2] == (0, kf.KF_T_BF16)
    # an explicit "quantizer" block next to the HF config wins over the vendor's
    both = {"hf_config": hf, "quantizer": {"mlp": {"quant_method": "RTN", "bits": 4}}}
    assert _quant_of(both, "model.layers.0.mlp.up_proj.weight")[1] == kf.KF_T_Q4
    assert _quant_of(both, "model.layers.0.self_attn.q_proj.weight")[1] == kf.KF_T_BF16


def test_awq_shard_windows_dequantise_to_the_windows_of_the_full_weight():
    # the tensor-parallel plan on the vendor AWQ layout (what kf_model_set_tensor_awq uploads on each rank): Q/K/V/gate/up keep a column range of
    # qweight / qzeros / scales, O / down a row range in whole 128-row groups.  Checked through the oracle's CU_Q42X_awq port: the window's arrays
    # dequantise to the window of the full dequantised weight, for every rank, and world 1 is the identity.
    import oracle_lib as ol
    lib = kf.load()
    hf = {"hidden_size": 512, "intermediate_size": 1024, "num_hidden_layers": 1, "num_attention_heads": 8, "num_key_value_heads": 4, "head_dim": 64,
          "vocab_size": 1024, "quantization_config": {"bits": 4, "group_size": 128, "quant_method": "awq", "zero_point": True}}
    text = json.dumps(hf).encode()
    for name, OC, IC in (("model.layers.0.self_attn.q_proj.weight", 512, 512), ("model.layers.0.self_attn.k_proj.weight", 256, 512),
                         ("model.layers.0.self_attn.o_proj.weight", 512, 512), ("model.layers.0.mlp.up_proj.weight", 1024, 512),
                         ("model.layers.0.mlp.down_proj.weight", 512, 1024)):
        w_io = ol.fill_normal(IC * OC, IC + OC + len(name), 0.05).reshape(IC, OC)
        qw, qz, sc = ol.awq_pack(w_io, IC, OC)
        full = ol.awq_dequant(qw, qz, sc, IC, OC)  # bf16 [in][out]
        for world in (1, 2, 4):
            for rank in range(world):
                shape = (C.c_int * 6)()
                assert lib.kf_config_shard_of(text, name.encode(), rank, world, shape, None) == 0
                _, _, OCl, ICl, r0, c0 = list(shape)
                n, err = C.c_size_t(0), C.c_void_p()
                assert lib.kf_config_awq_shard(text, name.encode(), rank, world, None, None, None, None, 0, C.byref(n), C.byref(err)) == 0
                assert n.value == ICl * OCl // 2 + (ICl // 128) * (OCl // 8) * 4 + (ICl // 128) * OCl * 2
                blob = np.zeros(n.value, dtype=np.uint8)
                assert lib.kf_config_awq_shard(text, name.encode(), rank, world, qw.ctypes.data, qz.ctypes.data, sc.ctypes.data, blob.ctypes.data,
                                               blob.nbytes, C.byref(n), C.byref(err)) == 0
                a, b = ICl * OCl // 2, ICl * OCl // 2 + (ICl // 128) * (OCl // 8) * 4
                lqw, lqz, lsc = blob[:a].view(np.uint32), blob[a:b].view(np.uint32), blob[b:].view(np.uint16)
                got = ol.awq_dequant(lqw, lqz, lsc, ICl, OCl)
                assert np.array_equal(got, full[c0:c0 + ICl, r0:r0 + OCl]), (name, world, rank)
                if world == 1:
                    assert np.array_equal(lqw, qw) and np.array_equal(lqz, qz) and np.array_equal(lsc, sc)
    # a window that would split a 128-row group or an int32 word is refused
    n, err = C.c_size_t(0), C.c_void_p()
    assert lib.kf_config_awq_shard(text, b"model.layers.0.self_attn.k_proj.weight", 0, 3, None, None, None, None, 0, C.byref(n), C.byref(err)) != 0
    if err.value:
        lib.kf_string_free(err)


def test_awq_repack_into_packedq_storage_matches_the_vendor_read():
    # gpt.awq_repack = 1: the window of an AWQ linear re-laid-out into PackedQ 4-bit words + gama (what kf_model_set_tensor_awq uploads then).  The
    # oracle's CU_Q128toX_ port reads it back: with unit scales the weights are the code differences q - z, so equality with CU_Q42X_awq's read of
    # the vendor arrays is bit-exact and pins the whole re-layout (nibble order, [in][out] -> [out][in], the 128-bit word layout, the group index);
    # with real scales the two reads differ only by the bf16 rounding of step = scale and zero = zero_point * scale.
    import oracle_lib as ol
    lib = kf.load()
    hf = {"hf_config": {"hidden_size": 512, "intermediate_size": 1024, "num_hidden_layers": 1, "num_attention_heads": 8, "num_key_value_heads": 4,
                        "head_dim": 64, "vocab_size": 1024,
                        "quantization_config": {"bits": 4, "group_size": 128, "quant_method": "awq", "zero_point": True}},
          "gpt": {"awq_repack": 1}}
    text = json.dumps(hf).encode()
    rng = np.random.default_rng(11)
    for name, OC, IC in (("model.layers.0.self_attn.q_proj.weight", 512, 512), ("model.layers.0.self_attn.o_proj.weight", 512, 512),
                         ("model.layers.0.mlp.up_proj.weight", 1024, 512), ("model.layers.0.mlp.down_proj.weight", 512, 1024)):
        for unit in (True, False):
            qw = rng.integers(0, 2 ** 32, size=IC * OC // 8, dtype=np.uint64).astype(np.uint32)
            qz = rng.integers(0, 2 ** 32, size=IC // 128 * OC // 8, dtype=np.uint64).astype(np.uint32)
            sc = (np.ones(IC // 128 * OC) if unit else rng.uniform(0.001, 0.02, size=IC // 128 * OC)).astype(np.float16)
            full = ol.awq_dequant(qw, qz, sc.view(np.uint16), IC, OC)  # bf16 [in][out], CU_Q42X_awq
            for world in (1, 2):
                for rank in range(world):
                    shape = (C.c_int * 6)()
                    assert lib.kf_config_shard_of(text, name.encode(), rank, world, shape, None) == 0
                    _, _, OCl, ICl, r0, c0 = list(shape)
                    n, err = C.c_size_t(0), C.c_void_p()
                    assert lib.kf_config_awq_shard(text, name.encode(), rank, world, None, None, None, None, 0, C.byref(n), C.byref(err)) == 0
                    nG = OCl * ICl // 128
                    assert n.value == OCl * ICl // 2 + 2 * (OCl + ICl + 2 * nG)  # szData + szGama of a Q4 tensor (GeQuant.cpp:518)
                    blob = np.zeros(n.value, dtype=np.uint8)
                    assert lib.kf_config_awq_shard(text, name.encode(), rank, world, qw.ctypes.data, qz.ctypes.data, sc.ctypes.data, blob.ctypes.data,
                                                   blob.nbytes, C.byref(n), C.byref(err)) == 0
                    data, gama = blob[:OCl * ICl // 2], blob[OCl * ICl // 2:].view(np.uint16)
                    got = ol.bf16_to_f32(ol.dequant(data, gama, OCl, ICl, 4, 128, 0))  # [out][in], CU_Q128toX_
                    want = ol.bf16_to_f32(full)[c0:c0 + ICl, r0:r0 + OCl].T
                    if unit:
                        assert np.array_equal(got, want), (name, world, rank)
                    else:
                        step = np.repeat(sc.astype(np.float32).reshape(IC // 128, OC)[c0 // 128:(c0 + ICl) // 128, r0:r0 + OCl].T, 128, axis=1)
                        # |step * q - zero| terms are each rounded to bf16 (2^-9 relative at most, q and z <= 15) plus the final rounding
                        assert np.all(np.abs(got - want) <= step * 15 * 3 * 2.0 ** -8), (name, world, rank)
                        assert np.abs(got - want).mean() <= 0.02 * step.mean()


def test_hf_quantization_config_mapping_equals_the_reference_function():
    # QUANT_CARD::Vendor2JSONx (CLI_params.cpp:240-262) compiled from the reference tree (oracle/_ref/libkoifish_refcpu.so) against the quantizer block
    # this library derives from the same HF config
    import oracle_lib as ol
    if ol.refcpu() is None:
        pytest.skip("oracle/_ref/libkoifish_refcpu.so not built (reference tree absent at build time)")
    lib = kf.load()
    for vendor in ({"bits": 4, "group_size": 128, "modules_to_not_convert": None, "quant_method": "awq", "version": "gemm", "zero_point": True},
                   {"bits": 4, "group_size": 128, "quant_method": "gptq", "desc_act": False, "sym": True},
                   {"quant_method": "awq", "bits": 4}):
        hf = {"hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 2, "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128,
              "quantization_config": vendor}
        out, err = C.c_void_p(), C.c_void_p()
        assert lib.kf_config_quantizer_json(json.dumps(hf).encode(), C.byref(out), C.byref(err)) == 0
        got = json.loads(C.cast(out, C.c_char_p).value.decode())
        lib.kf_string_free(out)
        want = ol.refcpu_vendor2jsonx(vendor)
        assert got == want and list(got) == list(want), (got, want)


def test_quantizer_card_selection_equals_the_reference_init4neuron():
    # QUANT_CARD::Init4Neuron (GeQuant.cpp:1186-1285) compiled from the reference tree, field by field against this library's card, over the
    # reference's own quantizer block (cases/qwen3/qwen3_596M_q4.json), the HF vendor mapping, and blocks exercising every branch
    import oracle_lib as ol
    if ol.refcpu() is None:
        pytest.skip("oracle/_ref/libkoifish_refcpu.so not built (reference tree absent at build time)")
    lib = kf.load()
    vendor = ol.refcpu_vendor2jsonx({"bits": 4, "group_size": 128, "quant_method": "awq", "zero_point": True, "version": "gemm"})
    blocks = [REF_Q4_QUANTIZER, vendor,
              {"group_size": 64, "self_attn": {"quant_method": "RTN", "bits": 2}, "mlp": {"quant_method": "yyang", "bits": 1}, "embed_tokens": {"bits": 8}},
              {"self_attn": {"bits": 4}, "mlp": {"quant_method": "yyang", "bits": 2, "group_size": 256}, "# lm_head": {"bits": 4}, "debug": {"x": 1}},
              {"q_proj": {"quant_method": "RTN", "bits": 4, "zero_point": True}, "down_proj": {"bits": 8}, "layers.1.": {"quant_method": "rtn", "bits": 2}},
              {"mlp": {"quant_method": "AWQ", "bits": 4}, "self_attn": {"bits": 3}, "filter": {"bits": 4, "filterQ": ["x"]}}]
    names = ["model.layers.0.self_attn.q_proj.weight", "model.layers.1.self_attn.o_proj.weight", "model.layers.0.mlp.up_proj.weight",
             "model.layers.1.mlp.down_proj.weight", "model.embed_tokens.weight", "lm_head.weight", "model.layers.0.input_layernorm.weight"]
    for q in blocks:
        cfg = kf.qwen3_config(2, 1024, 3072, 16, 8, quantizer=q)
        for name in names:
            out, errq, err = (C.c_int * 8)(), C.c_float(0), C.c_void_p()
            st = lib.kf_config_quant_card(json.dumps(cfg).encode(), name.encode(), out, C.byref(errq), C.byref(err))
            assert st == 0, name
            want, want_errq = ol.refcpu_init4neuron(name, q)
            assert list(out) == want and abs(errq.value - want_errq) < 1e-6, (q, name, list(out), want)


def test_safetensors_index_equals_the_reference_parser(tmp_path):
    # the reference's own safetensors reader (K_SafeTensors::MMAP -> mmap_from_file, src/Tensor/Safetensors.cpp, compiled into
    # oracle/_ref/libkoifish_refkun.so) on an HF-style file, against kf_safetensors_index: same tensors in file order, same dtypes, shapes and byte ranges
    import oracle_lib as ol
    from st_util import write_safetensors
    rng = np.random.default_rng(1)
    ts = [("model.norm.weight", "BF16", rng.integers(0, 65536, 64, dtype=np.uint16)),
          ("model.layers.0.mlp.up_proj.weight", "F16", rng.standard_normal((8, 64)).astype(np.float16)),
          ("lm_head.weight", "F32", rng.standard_normal((4, 64)).astype(np.float32)),
          ("model.layers.0.mlp.up_proj.qweight", "I32", rng.integers(0, 2 ** 31, (64, 1), dtype=np.int32))]
    p = tmp_path / "model.safetensors"
    write_safetensors(p, ts, metadata={"format": "pt"})
    got = ol.refkun_read(p)
    if got is None:
        pytest.skip("oracle/_ref/libkoifish_refkun.so not built (reference tree absent at build time)")
    mine = kf.safetensors_index(p)
    assert [(t["name"], t["dtype"], t["shape"], t["data_offsets"][1] - t["data_offsets"][0]) for t in got["tensors"]] == \
           [(e["name"], e["dtype"], e["shape"], e["nbytes"]) for e in mine]
    assert got["config"] is None


def _dims(text):
    lib = kf.load()
    info, err = kf.ModelInfo(), C.c_void_p()
    st = lib.kf_config_dims(text.encode(), C.byref(info), C.byref(err))
    msg = C.cast(err, C.c_char_p).value.decode() if err.value else ""
    if err.value:
        lib.kf_string_free(err)
    return st, info, msg


def test_config_parsing_koifish_and_hf():
    # the reference's own layout (cases/qwen3/qwen3_596M_q4.json), including '#'-comment keys and nested junk
    ref_like = {
        "version": "0.1.0", "quantizer": REF_Q4_QUANTIZER,
        "model": {"#hf-card": "/x/", "arch": "QWEN3", "parameter": {
            "Layer": 6, "transformer": {"Ctx": 1024, "Embed": 1024, "Ffn": 3072, "Head": 16, "KVHead": 8, "head_dim": 128},
            "tie_word_embeddings": True, "max_pos_embeddings": 32768},
            "backbone": {"embed_tokens": {"Embedding": []}, "layer": {"self_attn": {"QKV": []}, "mlp": {"FFN": []}}}},
        "train": {"batch": 16, "learning-rate": 0.0006}, "debug": {"prompts": ["hello", "天命"], "fake_quant": -1}, "seed": 42}
    st, info, msg = _dims(json.dumps(ref_like))
    assert st == 0, msg
    assert (info.n_layers, info.n_embd, info.n_ff, info.n_head, info.n_head_kv, info.head_dim) == (6, 1024, 3072, 16, 8, 128)
    assert info.vocab == 151936 and info.tie_word_embeddings == 1 and info.max_seq_len == 1024
    assert abs(info.rope_theta - 10000.0) < 1e-3 and abs(info.norm_rms_eps - 1e-6) < 1e-12
    hf = {"architectures": ["Qwen3ForCausalLM"], "hidden_size": 5120, "intermediate_size": 25600, "num_hidden_layers": 64,
          "num_attention_heads": 64, "num_key_value_heads": 8, "head_dim": 128, "vocab_size": 151936, "rope_theta": 1000000,
          "rms_norm_eps": 1e-06, "tie_word_embeddings": False, "max_position_embeddings": 40960, "model_type": "qwen3"}
    st, info, msg = _dims(json.dumps(hf))
    assert st == 0, msg
    assert (info.n_layers, info.n_embd, info.n_ff, info.n_head, info.n_head_kv) == (64, 5120, 25600, 64, 8)
    assert info.rope_theta == 1e6 and info.tie_word_embeddings == 0


@pytest.mark.parametrize("bad", ['{"model":', '{"model": {"arch": "GPT2", "parameter": {"Layer": 2}}}',
                                 '{"model": {"arch": "QWEN3", "parameter": {"Layer": 2}}}', "[]", ""])
def test_bad_configs_return_errors_not_exits(bad):
    st, _, msg = _dims(bad)
    assert st == kf.KF_ERR_BAD_ARG and msg


def test_integration_doc_names_real_symbols():
    p = os.path.join(ROOT, "INTEGRATION.md")
    if not os.path.exists(p):
        pytest.skip("INTEGRATION.md not written yet")
    used = set(re.findall(r"\b(kf_[a-z0-9_]+)\s*\(", open(p).read()))
    assert used and used <= _declared_symbols(), used - _declared_symbols()


# ---------------------------------------------------------------------------------------------- HF safetensors header (host only)
def test_safetensors_index_and_malformed_files(tmp_path):
    import numpy as np
    from st_util import write_safetensors
    p = tmp_path / "model.safetensors"
    write_safetensors(p, [("model.norm.weight", "BF16", np.arange(8, dtype=np.uint16)),
                          ("model.layers.0.mlp.up_proj.weight", "F32", np.ones((4, 6), dtype=np.float32)),
                          ("model.layers.0.self_attn.q_proj.qweight", "I32", np.zeros((2, 3), dtype=np.int32))], metadata={"format": "pt"})
    idx = kf.safetensors_index(p)
    assert [e["name"] for e in idx] == ["model.norm.weight", "model.layers.0.mlp.up_proj.weight", "model.layers.0.self_attn.q_proj.qweight"]
    assert idx[0] == {"name": "model.norm.weight", "dtype": "BF16", "shape": [8], "nbytes": 16}
    assert idx[1]["shape"] == [4, 6] and idx[1]["nbytes"] == 96 and idx[2]["dtype"] == "I32"
    # malformed: missing file, truncated, header length beyond the file, offsets that do not match shape x dtype
    with pytest.raises(kf.KoifishError):
        kf.safetensors_index(tmp_path / "nope.safetensors")
    raw = p.read_bytes()
    (tmp_path / "short.safetensors").write_bytes(raw[:5])
    (tmp_path / "cut.safetensors").write_bytes(raw[:40])
    (tmp_path / "len.safetensors").write_bytes((10 ** 9).to_bytes(8, "little") + raw[8:])
    (tmp_path / "off.safetensors").write_bytes(raw.replace(b'"data_offsets":[0,16]', b'"data_offsets":[0,12]'))
    for name in ("short", "cut", "len", "off"):
        with pytest.raises(kf.KoifishError):
            kf.safetensors_index(tmp_path / (name + ".safetensors"))


def test_safetensors_dtype_conversions_match_the_oracle(tmp_path):
    # F32 and F16 sources are rounded to bf16 (nearest even) exactly as the oracle's f32 -> bf16; incl. fp16 subnormals, infinities, signed zeros
    import numpy as np
    import oracle_lib as ol
    from st_util import write_safetensors
    rng = np.random.default_rng(1)
    f32 = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 10.0 ** rng.integers(-8, 8, 4096),
                          np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 1.00390625, 1.01171875, 3.3895314e38, 1e-40], dtype=np.float32)])
    f16 = np.concatenate([rng.standard_normal(4096).astype(np.float16), np.array([0.0, -0.0, 6e-8, -6e-8, 6.1e-5, 65504.0, np.inf, -np.inf], dtype=np.float16)])
    b16 = rng.integers(0, 65536, size=512).astype(np.uint16)
    p = tmp_path / "t.safetensors"
    write_safetensors(p, [("a", "F32", f32), ("b", "F16", f16), ("c", "BF16", b16)])
    assert np.array_equal(kf.safetensors_read_bf16(p, "a", f32.size), ol.f32_to_bf16(f32))
    assert np.array_equal(kf.safetensors_read_bf16(p, "b", f16.size), ol.f32_to_bf16(f16.astype(np.float32)))
    assert np.array_equal(kf.safetensors_read_bf16(p, "c", b16.size), b16)
    with pytest.raises(kf.KoifishError):
        kf.safetensors_read_bf16(p, "nope", 8)
    with pytest.raises(kf.KoifishError):
        kf.safetensors_read_bf16(p, "a", 8)  # buffer too small


def test_bench_reference_arm_under_torchrun_two_ranks():
    # the driver launches `--impl reference` like our own arm (torchrun for N > 1): rank 0 alone runs the reference's CPU primitives and
    # prints ONE json line describing our arm's workload; the other rank exits 0 without output.  Exactly --steps timed samples.
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["metric"].startswith("decode tokens/s") and d["unit"] == "tokens/s" and d["value"] > 0
    assert d["config"]["workload"] == "Qwen3-32B decode, batch 1, ctx 512, qwen3-32b-q4" and d["config"]["global_batch"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    if d["cpu_baseline"]["kind"] == "reference":  # oracle/_ref present: the reference's own primitives, exactly --steps timed samples
        assert "3 timed" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
