"""A fish.kun file written by the reference's OWN writer -- K_SafeTensors::Register + insertJS + Save driven as Fish::SAFETENSOR_Serialize drives them
(reference src/Manifold/Serialize.cpp:286-360, 554-680, 860-960; src/Tensor/Safetensors.hpp:87-102), compiled from the reference tree into
oracle/_ref/libkoifish_refkun.so (oracle/ref_kun.cpp) -- on seeded host payloads.  Writes tests/golden/ref_written.kun (a few KB); the CPU suite reads
it with this library's reader without needing /root/reference.
Run in the build container:  python tests/golden/make_golden_refkun.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CONFIG = {"vendor": "gruai", "CLI_params": {"config": {"model": {"arch": "QWEN3", "parameter": {"Layer": 2, "transformer": {"Embed": 128, "Ffn": 256}}},
                                                       "quantizer": {"mlp": {"bits": 4, "quant_method": "RTN"}}, "seed": 42, "lr": 6.0e-4}},
          "tokenizer": {"tokens": ""}, "tensors": {"model.norm.weight": 0}}


def tensors():
    rng = np.random.default_rng(20261018)
    out = []
    for name, dt, shape, bits in (("model.layers.0.mlp.up_proj.weight", "Q<4>", (256, 128), 4), ("model.layers.0.mlp.down_proj.weight", "TERNARY", (128, 256), 2),
                                  ("model.layers.0.self_attn.q_proj.weight", "BINARY", (128, 128), 1), ("model.layers.0.self_attn.k_proj.weight", "F8E5M2", (64, 128), 8),
                                  ("model.embed_tokens.weight", "BF16(E8)", (32, 128), 16), ("model.norm.weight", "BF16(E8)", (128,), 16)):
        numel = int(np.prod(shape))
        szd = numel * bits // 8
        szg = 2 * (shape[0] + shape[1] + 2 * (numel // 128)) if bits < 8 else 0
        out.append((name, dt, shape, szd, szg, rng.integers(0, 256, szd + szg, dtype=np.uint8).tobytes()))
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "ref_written.kun")
    assert ol.refkun_write(path, CONFIG, tensors()), "oracle/_ref/libkoifish_refkun.so is not built (needs /root/reference)"
    print("wrote", path, os.path.getsize(path), "bytes")
