"""CPU tests of the fish.kun container code (csrc/Tensor/KunFile.cpp) through the host-only C ABI entries: files laid out the
way the reference's writer lays them out (K_SafeTensors::Save / GTensor::jDesc / insertJS, reference src/Manifold/Serialize.cpp:61-100, 880-1010,
src/Tensor/Safetensors.hpp:87-102) are read back, files this library writes are checked byte for byte with plain Python, and the msgpack config
entry goes both ways against the `msgpack` package."""
import json
import random

import numpy as np
import pytest

import koifish_b200 as kf
from st_util import write_kun_reference_style as write_reference_style

msgpack = pytest.importorskip("msgpack")


def sample_tensors(rng):
    out = []
    for name, dt, shape, bits in (("model.layers.0.mlp.up_proj.weight", "Q<4>", (512, 256), 4), ("model.layers.0.mlp.down_proj.weight", "TERNARY", (256, 512), 2),
                                  ("model.layers.0.self_attn.q_proj.weight", "BINARY", (256, 256), 1), ("model.layers.0.self_attn.k_proj.weight", "F8E5M2", (128, 256), 8),
                                  ("model.embed_tokens.weight", "BF16(E8)", (1024, 256), 16), ("model.norm.weight", "BF16(E8)", (256,), 16)):
        numel = int(np.prod(shape))
        szd = numel * bits // 8
        szg = 2 * (shape[0] + shape[1] + 2 * (numel // 128)) if bits < 8 else 0  # GeQuant.cpp:518
        out.append((name, dt, shape, szd, szg, rng.integers(0, 256, szd + szg, dtype=np.uint8).tobytes()))
    return out


CONFIG = {"vendor": "gruai", "CLI_params": {"config": {"model": {"arch": "QWEN3", "parameter": {"Layer": 2, "transformer": {"Embed": 256, "Ffn": 512}}},
                                                        "quantizer": {"mlp": {"bits": 4, "quant_method": "RTN"}}, "seed": 42, "lr": 6.0e-4}},
          "tokenizer": {"tokens": ""}, "tensors": {"model.norm.weight": 0, "big": 5000000000, "neg": -70000}}


def test_reads_a_file_laid_out_like_the_reference_writer(tmp_path):
    rng = np.random.default_rng(1)
    tensors = sample_tensors(rng)
    p = tmp_path / "fish.kun"
    write_reference_style(p, tensors, CONFIG)
    idx = kf.kun_index(p)
    assert [e["name"] for e in idx] == [t[0] for t in tensors]  # file order; the config entry is not a tensor
    off = 0
    for e, (name, dt, shape, szd, szg, _) in zip(idx, tensors):
        assert (e["dtype"], tuple(e["shape"]), e["szData"], e["szGama"], e["offset"]) == (dt, tuple(shape), szd, szg, off)
        off += szd + szg
    assert kf.kun_config(p) == CONFIG
    write_reference_style(p, tensors, None)
    assert kf.kun_config(p) is None and len(kf.kun_index(p)) == len(tensors)


def test_writer_output_checked_with_plain_python(tmp_path):
    rng = np.random.default_rng(2)
    tensors = sample_tensors(rng)
    p = tmp_path / "out.kun"
    kf.kun_write(p, CONFIG, tensors)
    raw = open(p, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    header = json.loads(raw[8:8 + n])
    data = raw[8 + n:]
    assert header["__metadata__"] == {"format": "pt", "writer": "koifish"}  # K_SafeTensors::UpdateMetaData
    assert list(header)[1:-1] == [t[0] for t in tensors] and list(header)[-1] == "__koifish__config__"
    off = 0
    for name, dt, shape, szd, szg, blob in tensors:
        e = header[name]
        assert e == {"dtype": dt, "shape": list(shape), "data_offsets": [off, off + szd + szg], "loAB": 0, "szGama": szg, "szData": szd}  # GTensor::jDesc
        assert data[off:off + szd + szg] == blob
        off += szd + szg
    c = header["__koifish__config__"]
    assert c["dtype"] == "U8" and c["data_offsets"] == [off, len(data)] and c["shape"] == [len(data) - off]
    assert msgpack.unpackb(data[off:], raw=False) == CONFIG
    # and back through the reader
    assert kf.kun_config(p) == CONFIG
    assert [(e["name"], e["offset"]) for e in kf.kun_index(p)] == [(t[0], header[t[0]]["data_offsets"][0]) for t in tensors]


def test_malformed_and_out_of_scope_files_are_refused(tmp_path):
    rng = np.random.default_rng(3)
    tensors = sample_tensors(rng)
    p = tmp_path / "bad.kun"
    write_reference_style(p, tensors, CONFIG, moments=3)  # a training-state checkpoint: weights + two optimizer moments per payload
    with pytest.raises(kf.KoifishError, match="optimizer state"):
        kf.kun_index(p)
    name, dt, shape, szd, szg, blob = tensors[0]
    for bad in ((name, "Q<5>", shape, szd, szg, blob), (name, dt, shape, szd + 16, szg, blob + bytes(16)), (name, dt, (511, 256), szd, szg, blob),
                (name, dt, shape, szd, szg + 2, blob)):
        write_reference_style(p, [bad], None)
        with pytest.raises(kf.KoifishError):
            kf.kun_index(p)
    with pytest.raises(kf.KoifishError):
        kf.kun_write(p, None, [(name, dt, shape, szd - 1, szg, blob)])
    with pytest.raises(kf.KoifishError):
        kf.kun_write(p, None, [tensors[0], tensors[0]])
    open(p, "wb").write(b"\x05\x00")
    with pytest.raises(kf.KoifishError):
        kf.kun_index(p)
    open(p, "wb").write((1 << 40).to_bytes(8, "little") + b"{}")
    with pytest.raises(kf.KoifishError):
        kf.kun_index(p)
    with pytest.raises(kf.KoifishError):
        kf.kun_index(tmp_path / "missing.kun")


def _random_json(rng, depth=0):
    k = rng.randint(0, 9 if depth < 4 else 6)
    if k == 0:
        return None
    if k == 1:
        return rng.random() < 0.5
    if k == 2:
        return rng.choice([0, 1, -1, 31, -32, -33, 127, 128, 255, 256, -128, -129, 65535, 65536, -32768, -32769, 2 ** 31 - 1, 2 ** 31, -2 ** 31, -2 ** 31 - 1,
                           2 ** 32, 2 ** 40, -2 ** 40, 2 ** 53 - 1, -(2 ** 53 - 1), rng.randint(-10 ** 6, 10 ** 6)])
    if k == 3:
        return rng.choice([0.5, -1.25, 3.141592653589793, 1e-9, 6.0e-4, 1e300, -2.5e-300, 1.0000001])
    if k in (4, 5):
        n = rng.choice([0, 1, 5, 31, 32, 33, 200, 255, 256, 300])
        return "".join(rng.choice("ab c\"\\\n\té天😀") for _ in range(n))
    if k == 6:
        return "x" * rng.choice([65535, 65536, 70000])
    if k in (7, 8):
        return [_random_json(rng, depth + 1) for _ in range(rng.choice([0, 1, 3, 15, 16, 17, 40]))]
    return {("k%d_%s" % (i, rng.choice(["", "é", "天"]))): _random_json(rng, depth + 1) for i in range(rng.choice([0, 1, 3, 15, 16, 17, 40]))}


def test_msgpack_config_both_ways_against_the_msgpack_package(tmp_path):
    rng = random.Random(5)
    p = tmp_path / "cfg.kun"
    for i in range(60):
        cfg = {"case": i, "value": _random_json(rng)}
        write_reference_style(p, [], cfg)          # packed by the msgpack package, decoded here
        assert kf.kun_config(p) == cfg, i
        kf.kun_write(p, cfg, [])                   # encoded here, unpacked by the msgpack package
        raw = open(p, "rb").read()
        n = int.from_bytes(raw[:8], "little")
        assert msgpack.unpackb(raw[8 + n:], raw=False, strict_map_key=False) == cfg, i
    # floats the reference's writer (nlohmann::json::to_msgpack) stores as float32 when exact
    write_reference_style(p, [], {"a": 0.5})
    raw = bytearray(open(p, "rb").read())
    packed32 = msgpack.packb({"a": 0.5}, use_single_float=True)
    n = int.from_bytes(raw[:8], "little")
    header = json.loads(raw[8:8 + n])
    header["__koifish__config__"]["shape"] = [len(packed32)]
    header["__koifish__config__"]["data_offsets"] = [0, len(packed32)]
    text = json.dumps(header).encode()
    open(p, "wb").write(len(text).to_bytes(8, "little") + text + packed32)
    assert kf.kun_config(p) == {"a": 0.5}


def test_files_written_here_are_read_by_the_reference_reader(tmp_path):
    # the reference's own checkpoint reader compiled from its tree (oracle/ref_kun.cpp -> oracle/_ref/libkoifish_refkun.so): K_SafeTensors::MMAP
    # (mmap_from_file + validate_data_offsets + loadJS of the msgpack config; Serialize.cpp:428-494) on a file kf_kun_write produced, then its own
    # GTensor::jDesc of every tensor it parsed.  The reference must find the same tensors, in the same order, with the header entries this library
    # wrote, and decode the same config.
    import oracle_lib as ol
    rng = np.random.default_rng(9)
    tensors = sample_tensors(rng)
    p = tmp_path / "mine.kun"
    kf.kun_write(p, CONFIG, tensors)
    got = ol.refkun_read(p)
    if got is None:
        pytest.skip("oracle/_ref/libkoifish_refkun.so not built (reference tree absent at build time)")
    assert got["config"] == CONFIG
    parsed = [t for t in got["tensors"] if t["name"] != "__koifish__config__"]
    assert [t["name"] for t in parsed] == [t[0] for t in tensors]
    off = 0
    for t, (name, dt, shape, szd, szg, _) in zip(parsed, tensors):
        assert t == {"dtype": dt, "shape": list(shape), "data_offsets": [off, off + szd + szg], "loAB": 0, "szGama": szg, "szData": szd, "name": name}
        off += szd + szg
    mine = {e["name"]: e for e in kf.kun_index(p)}  # and this library's reader agrees with the reference's on the same file
    for t in parsed:
        e = mine[t["name"]]
        assert (e["dtype"], e["shape"], e["szData"], e["szGama"], e["offset"]) == (t["dtype"], t["shape"], t["szData"], t["szGama"], t["data_offsets"][0])


def _check_reference_written(path, tensors, config):
    idx = kf.kun_index(path)
    assert [e["name"] for e in idx] == [t[0] for t in tensors]
    raw = open(path, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    data = raw[8 + n:]
    off = 0
    for e, (name, dt, shape, szd, szg, blob) in zip(idx, tensors):
        assert (e["dtype"], tuple(e["shape"]), e["szData"], e["szGama"], e["offset"]) == (dt, tuple(shape), szd, szg, off), name
        assert data[off:off + szd + szg] == bytes(blob), name
        off += szd + szg
    assert kf.kun_config(path) == config


def test_reads_the_golden_file_written_by_the_reference_writer():
    # tests/golden/ref_written.kun: produced by the reference's own K_SafeTensors::Register / insertJS / Save (tests/golden/make_golden_refkun.py)
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden_refkun as g
    _check_reference_written(os.path.join(here, "ref_written.kun"), g.tensors(), g.CONFIG)


def test_reads_files_written_by_the_reference_writer_live(tmp_path):
    import oracle_lib as ol
    rng = np.random.default_rng(21)
    tensors = sample_tensors(rng)
    p = tmp_path / "ref.kun"
    if not ol.refkun_write(p, CONFIG, tensors):
        pytest.skip("oracle/_ref/libkoifish_refkun.so not built (reference tree absent at build time)")
    _check_reference_written(p, tensors, CONFIG)
    # and the two writers agree byte for byte on everything but the JSON header's spacing / padding
    q = tmp_path / "mine.kun"
    kf.kun_write(q, CONFIG, tensors)
    a, b = open(p, "rb").read(), open(q, "rb").read()
    na, nb = int.from_bytes(a[:8], "little"), int.from_bytes(b[:8], "little")
    assert json.loads(a[8:8 + na]) == json.loads(b[8:8 + nb])
    assert list(json.loads(a[8:8 + na])) == list(json.loads(b[8:8 + nb]))  # same key order
    assert a[8 + na:] == b[8 + nb:]
