"""GPU tests of the fish.kun container at the model level: kf_model_save_kun / kf_model_load_kun -- what the reference's
Fish::SAFETENSOR_Serialize writes and SAFETENSOR2Gensors + GTensor::LoadParam + Serial_Quant_MMAP read (reference src/Manifold/Serialize.cpp:
145-230, 770-1010; src/Device/CUDA/huTensor.cu:413-458, 487-588).  The container code itself is covered on the CPU (tests/test_kun_host.py)."""
import json

import numpy as np
import pytest

import koifish_b200 as kf
from st_util import write_kun_reference_style

pytestmark = pytest.mark.gpu

QUANTIZER = {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "yyang", "bits": 2}, "embed_tokens": {"bits": 8}}


@pytest.fixture(scope="module")
def ctx():
    c = kf.Context(0)
    yield c
    c.close()


def _cfg(quantizer=QUANTIZER):
    return kf.qwen3_config(2, 256, 512, 4, 2, 64, 1024, quantizer, False, 64, 1, 42, 1e6)


def _same_logits(a, b, n=4):
    for pos in range(n):
        tok = (1000 + 37 * pos) % 1024
        la, _ = a.forward([tok], [pos])
        lb, _ = b.forward([tok], [pos])
        assert np.array_equal(la, lb), pos


def test_kun_round_trip_and_a_reference_style_file(ctx, tmp_path):
    a = kf.Model(ctx, _cfg())
    a.init_random()
    p = tmp_path / "fish.kun"
    a.save_kun(p)
    names = a.tensor_names()
    idx = {e["name"]: e for e in kf.kun_index(p)}
    assert sorted(idx) == sorted(names)
    for name, e in idx.items():
        d = a.tensor_desc(name)
        want = ("TERNARY" if "mlp" in name else "Q<4>" if "self_attn" in name and "proj" in name else "F8E5M2" if "embed_tokens" in name else "BF16(E8)")
        assert e["dtype"] == want, name
        assert e["shape"] == ([d.cols] if d.rows == 1 else [d.rows, d.cols]), name
        bits = {"TERNARY": 2, "Q<4>": 4, "F8E5M2": 8, "BF16(E8)": 16}[want]
        assert e["szData"] == d.rows * d.cols * bits // 8, name
        assert e["szGama"] == (2 * (d.rows + d.cols + 2 * (d.rows * d.cols // 128)) if bits < 8 else 0), name  # szGama, GeQuant.cpp:518
    cfg = kf.kun_config(p)
    assert cfg["vendor"] == "koifish_b200" and cfg["CLI_params"]["config"] == json.loads(json.dumps(_cfg()))
    b = kf.Model(ctx, _cfg())
    assert b.load_kun(p) == (len(names), 0)
    for name in ("model.layers.1.mlp.down_proj.weight", "model.layers.0.self_attn.q_proj.weight", "model.embed_tokens.weight", "lm_head.weight",
                 "model.layers.0.self_attn.q_norm.weight"):
        assert np.array_equal(a.dequant_tensor(name), b.dequant_tensor(name)), name
    _same_logits(a, b)
    # the same payloads in a file laid out by plain Python the way the reference's writer does: other key order, an entry the model does not have
    raw = open(p, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    header, data = json.loads(raw[8:8 + n]), raw[8 + n:]
    tensors = []
    for name in reversed(names):
        e = header[name]
        tensors.append((name, e["dtype"], e["shape"], e["szData"], e["szGama"], data[e["data_offsets"][0]:e["data_offsets"][1]]))
    tensors.append(("model.out.weight", "BF16(E8)", [8, 8], 128, 0, bytes(128)))
    q = tmp_path / "ref_style.kun"
    write_kun_reference_style(q, tensors, {"vendor": "gruai", "CLI_params": {"config": {}}})
    c = kf.Model(ctx, _cfg())
    assert c.load_kun(q) == (len(names), 1)
    _same_logits(a, c)


def test_kun_of_another_configuration_is_refused(ctx, tmp_path):
    a = kf.Model(ctx, _cfg())
    a.init_random()
    p = tmp_path / "fish.kun"
    a.save_kun(p)
    other = dict(QUANTIZER, mlp={"quant_method": "RTN", "bits": 4})  # this config stores the mlp linears as Q<4>, the file holds TERNARY
    with pytest.raises(kf.KoifishError):
        kf.Model(ctx, _cfg(other)).load_kun(p)
    wider = kf.Model(ctx, kf.qwen3_config(2, 256, 1024, 4, 2, 64, 1024, QUANTIZER, False, 64, 1, 42, 1e6))  # another Ffn: shapes differ
    with pytest.raises(kf.KoifishError):
        wider.load_kun(p)
    with pytest.raises(kf.KoifishError):
        a.load_kun(tmp_path / "missing.kun")
    with pytest.raises(kf.KoifishError):
        kf.Model(ctx, _cfg()).save_kun(p)  # nothing resident yet
