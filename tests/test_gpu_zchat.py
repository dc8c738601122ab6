"""GPU tests of the chat loop: kf_model_generate -- the generation loop of Fish::Chat, reference src/Manifold/GoPT.cpp:1111-1235 --
against the same steps made one kf_model_forward call at a time, its three stop conditions, and one text-in / text-out turn through the tokenizer."""
import os

import numpy as np
import pytest

import koifish_b200 as kf

pytestmark = pytest.mark.gpu

Q4 = {"group_size": 128, "self_attn": {"quant_method": "RTN", "bits": 4}, "mlp": {"quant_method": "RTN", "bits": 4}}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer")


@pytest.fixture(scope="module")
def ctx():
    c = kf.Context(0)
    yield c
    c.close()


def _model(ctx, vocab=1024, max_seq=128):
    m = kf.Model(ctx, kf.qwen3_config(2, 256, 512, 4, 2, 64, vocab, Q4, True, max_seq, 1, 42, 1e6, norm_sigma=0.1))
    m.init_random()
    return m


def _by_hand(model, prompt, n_new, pos0=0, panel=64):
    """the steps of Fish::Generate as separate public calls: prefill panels (last token's argmax), then one token per forward"""
    nxt = None
    for off in range(0, len(prompt), panel):
        part = prompt[off:off + panel]
        last = off + len(part) == len(prompt)
        _, nxt = model.forward(part, list(range(pos0 + off, pos0 + off + len(part))), seq_mode=2, want_logits=False, want_next=last)
    out, pos = [], pos0 + len(prompt)
    tok = int(nxt[0])
    while len(out) < n_new:
        out.append(tok)
        if len(out) == n_new:
            break
        _, nxt = model.forward([tok], [pos], want_logits=False, want_next=True)
        tok, pos = int(nxt[0]), pos + 1
    return out


@pytest.mark.parametrize("n_prompt", [1, 5, 20, 70])  # one token, a short panel, a tensor-core panel, two panels (64 + 6)
def test_generate_equals_the_same_steps_by_hand(ctx, n_prompt):
    a, b = _model(ctx), _model(ctx)
    prompt = [(1000 + 37 * i) % 1024 for i in range(n_prompt)]
    want = _by_hand(b, prompt, 12)
    got, why = a.generate(prompt, 12)
    assert (got, why) == (want, 2)
    # a second turn continues the conversation from the cached context: the last generated token was never fed, so the turn starts with it
    turn2, pos0 = [got[-1], 7, 8, 9], n_prompt + 12 - 1
    got2, why = a.generate(turn2, 5, pos0=pos0)
    assert (got2, why) == (_by_hand(b, turn2, 5, pos0=pos0), 2)


def test_generate_stop_conditions(ctx):
    m = _model(ctx)
    prompt = [(1000 + 37 * i) % 1024 for i in range(70)]
    free, why = m.generate(prompt, 100)          # window 128: positions 70 .. 127 can still be fed
    assert why == 3 and len(free) == 128 - 70 + 1
    eos = free[3]
    k = free.index(eos)
    got, why = m.generate(prompt, 100, eos_id=eos)
    assert (got, why) == (free[:k], 1)            # the eos token itself is not emitted
    got, why = m.generate(prompt, 4)
    assert (got, why) == (free[:4], 2)
    assert m.generate(prompt, 0) == ([], 2)
    got, why = m.generate([1] * 128, 5)           # a prompt that fills the window: one token can be drawn, none fed
    assert why == 3 and len(got) == 1
    for bad in (dict(prompt_ids=[], max_new_tokens=4), dict(prompt_ids=[1] * 129, max_new_tokens=4), dict(prompt_ids=[1], max_new_tokens=4, pos0=128),
                dict(prompt_ids=[5000], max_new_tokens=4)):
        with pytest.raises(kf.KoifishError):
            m.generate(**bad)


def test_sampled_generation_is_seeded(ctx):
    m = _model(ctx)
    prompt = [(1000 + 37 * i) % 1024 for i in range(9)]

    def run(seed):
        m.set_sampler(0.9, 40, 0.95, seed)
        return m.generate(prompt, 16)[0]

    a = run(11)
    assert a == run(11) and a != run(12)
    m.set_sampler(0.0, 1, 1.0, 0)
    assert m.generate(prompt, 16)[0] == _by_hand(_model(ctx), prompt, 16)


def test_one_chat_turn_text_in_text_out(ctx):
    tok = kf.Tokenizer(GOLD)
    m = _model(ctx, vocab=tok.vocab_size)  # 1056 = 66 x 16
    text, ids, why = kf.chat_once(m, tok, "What is the capital of France?", "You are a helpful assistant.", max_new_tokens=24)
    prompt = tok.encode(kf.chatml_prompt("What is the capital of France?", "You are a helpful assistant."))
    want, why2 = _model(ctx, vocab=tok.vocab_size).generate(prompt, 24, tok.eos_id)
    assert (ids, why) == (want, why2)
    assert text == tok.decode(want, skip_special_tokens=True)
    assert tok.eos_id not in ids
