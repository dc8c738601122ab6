"""The reference's own integration tests of this path, cases/test_lite.py:36-52 (`bubble --hf ./Models/<name>/ --prompts "..."` + a substring of
the chat output), through this library: HF checkpoint directory -> kf.from_pretrained -> one ChatML turn -> text.  They need the real
checkpoints, which do not exist offline: every test skips unless KF_MODELS_DIR points at a directory holding Qwen3-0.6B / Qwen3-4B /
Qwen3-4B-AWQ (the reference's ./Models).  Sampler: the reference's CHAT_SAMPLER defaults (T 0.6, top-k 50, top-p 0.95, seed 42;
src/CLI_params.hpp:677-683) and, because a sampled answer depends on the generator, greedy as well -- either may satisfy the substring."""
import os

import pytest

import koifish_b200 as kf

pytestmark = pytest.mark.gpu

MODELS = os.environ.get("KF_MODELS_DIR", "")
SALLY = "Sally (a girl) has 3 brothers. Each brother has 2 sisters. How many sisters does Sally have?"


def _chat(name, prompt, quantizer=None, max_new=512):
    path = os.path.join(MODELS, name)
    if not MODELS or not os.path.isdir(path):
        pytest.skip("set KF_MODELS_DIR to a directory holding %s (HF checkpoint: config.json, *.safetensors, tokenizer.json)" % name)
    ctx = kf.Context(0)
    try:
        model, tok = kf.from_pretrained(ctx, path, quantizer=quantizer, max_seq_len=1024)
        outs = []
        for temperature in (0.0, 0.6):
            model.set_sampler(temperature, 50, 0.95, 42)
            text, _, _ = kf.chat_once(model, tok, prompt, max_new_tokens=max_new, enable_thinking=False)
            outs.append(text)
        return outs
    finally:
        ctx.close()


def _sally_ok(text):
    return any(s in text for s in ("Answer: \\boxed{1}", "Answer: 1", "Answer:1", "answer:1", "\\boxed{1}"))


def test_chat_qwen3_596M():  # cases/test_lite.py:36-38
    assert any("Hello! How can I assist you today?" in t for t in _chat("Qwen3-0.6B", "hello", max_new=64))


def test_chat_qwen3_4B():  # cases/test_lite.py:40-43
    assert any(_sally_ok(t) for t in _chat("Qwen3-4B", SALLY))


def test_chat_qwen3_4B_awq():  # cases/test_lite.py:50-52: the vendor-quantised checkpoint, read in its own layout
    assert any(_sally_ok(t) for t in _chat("Qwen3-4B-AWQ", SALLY))


def test_chat_qwen3_596M_quantised_at_load():  # the same checkpoint through the 4-bit RTN path of cases/qwen3/qwen3_596M_q4.json's quantizer block
    q = {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}
    assert any("Hello" in t for t in _chat("Qwen3-0.6B", "hello", quantizer=q, max_new=64))
