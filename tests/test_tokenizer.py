"""CPU tests of csrc/TokenSet through the C ABI (include/kf_tokenizer.h) against
  * golden vectors computed by the HF `tokenizers` library (tests/golden/make_tokenizer_golden.py -> tests/golden/tokenizer/),
  * CPython's unicodedata for Unicode NFC,
  * the `tokenizers` library itself, live, on random strings -- when it is importable (it is in this image).
Bit-exact: token ids, pre-tokenisation pieces and decoded text must equal the library's."""
import json
import os
import random
import unicodedata

import pytest

import koifish_b200 as kf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def tok():
    return kf.Tokenizer(GOLD)  # a directory: tokenizer.json + tokenizer_config.json (AutoTokenizer::from_pretrained)


@pytest.fixture(scope="module")
def tok_l3():
    return kf.Tokenizer(os.path.join(GOLD, "llama3style", "tokenizer.json"))


def test_golden_ids_pieces_and_decoding(tok, tok_l3, gold):
    assert tok.vocab_size == gold["vocab_size"]
    for c in gold["cases"]:
        text = c["text"]
        assert kf.nfc(text) == c["nfc"], repr(text)
        assert tok.pre_tokenize(c["nfc"]) == c["pieces"], repr(text)
        assert tok_l3.pre_tokenize(c["nfc"]) == c["pieces_llama3"], repr(text)
        ids = tok.encode(text)
        assert ids == c["ids"], repr(text)
        assert tok_l3.encode(text) == c["ids_llama3"], repr(text)
        assert tok.decode(ids) == c["decoded"], repr(text)
        assert tok.decode(ids, skip_special_tokens=True) == c["decoded_skip_special"], repr(text)


def test_special_tokens_and_eos(tok, gold):
    for s, i in gold["specials"].items():
        assert tok.token_to_id(s) == i and tok.id_to_token(i) == s
        assert tok.encode(s) == [i]
        assert tok.is_special(i) == s.startswith("<|")
    assert tok.eos_id == gold["specials"]["<|im_end|>"]        # tokenizer_config.json "eos_token"
    assert tok.pad_id == gold["specials"]["<|endoftext|>"]
    assert tok.bos_id == -1
    assert tok.token_to_id("no such token") == -1
    with pytest.raises(kf.KoifishError):
        tok.id_to_token(tok.vocab_size)
    # T2STR of single tokens (what Fish::Chat prints piece by piece): the pieces of a text concatenate back to it for ASCII
    text = "Hello, world! It's 42."
    assert "".join(tok.decode([i]) for i in tok.encode(text)) == text


def test_decoding_across_a_cut_multibyte_character_is_lossy_like_the_library(tok, gold):
    for c in gold["lossy"]:
        assert tok.decode(c["ids"]) == c["decoded"], c["ids"]


def test_invalid_utf8_and_unsupported_pipelines_fail_loudly(tok):
    with pytest.raises(kf.KoifishError):
        tok.encode(b"abc\xff\xfe")
    with pytest.raises(kf.KoifishError):
        tok.encode(b"\xe4\xb8")  # a truncated 3-byte character
    with open(os.path.join(GOLD, "tokenizer.json")) as f:
        j = json.load(f)
    for mutate in (lambda d: d.__setitem__("normalizer", {"type": "NFKC"}),
                   lambda d: d["pre_tokenizer"]["pretokenizers"][0]["pattern"].__setitem__("Regex", r"\w+|\s+"),
                   lambda d: d["pre_tokenizer"]["pretokenizers"][1].__setitem__("use_regex", True),
                   lambda d: d["model"].__setitem__("byte_fallback", True),
                   lambda d: d.__setitem__("pre_tokenizer", {"type": "Whitespace"}),
                   lambda d: d["model"].__setitem__("type", "WordPiece"),
                   lambda d: d["added_tokens"][0].__setitem__("lstrip", True),
                   lambda d: d["model"]["merges"].append(["zz", "nope"])):
        d = json.loads(json.dumps(j))
        mutate(d)
        with pytest.raises(kf.KoifishError):
            kf.Tokenizer(json_text=json.dumps(d))
    with pytest.raises(kf.KoifishError):
        kf.Tokenizer("/nonexistent/dir")
    # the same file from memory, merges in the older "a b" string form, no config: eos falls back to the family's usual name
    d = json.loads(json.dumps(j))
    d["model"]["merges"] = [" ".join(m) for m in d["model"]["merges"]]
    t2 = kf.Tokenizer(json_text=json.dumps(d))
    assert t2.encode("Hello world, it's 2024!") == tok.encode("Hello world, it's 2024!")
    assert t2.eos_id == tok.token_to_id("<|im_end|>")


def test_nfc_equals_unicodedata():
    # every code point on its own, then random sequences biased towards combining marks, Hangul jamo and composites
    rng = random.Random(7)
    singles = [chr(c) for c in range(1, 0x110000) if not 0xD800 <= c <= 0xDFFF]  # U+0000 cannot cross a NUL-terminated C string
    for i in range(0, len(singles), 4096):
        s = "a".join(singles[i:i + 4096])
        assert kf.nfc(s) == unicodedata.normalize("NFC", s), hex(i)
    marks = [chr(c) for c in range(0x110000) if unicodedata.combining(chr(c))]
    decomposable = [chr(c) for c in range(0x110000) if unicodedata.decomposition(chr(c)) and not unicodedata.decomposition(chr(c)).startswith("<")]
    jamo = [chr(c) for c in list(range(0x1100, 0x1113)) + list(range(0x1161, 0x1176)) + list(range(0x11A7, 0x11C3))] + [chr(0xAC00 + 28 * k) for k in range(40)]
    base = list("aeiouAEIOUnNcCyYΑαΙιاويकडডେෙဥ")
    for _ in range(3000):
        n = rng.randint(1, 12)
        s = "".join(rng.choice(rng.choice((marks, marks, decomposable, jamo, base, base))) for _ in range(n))
        assert kf.nfc(s) == unicodedata.normalize("NFC", s), [hex(ord(ch)) for ch in s]


def test_chatml_templates_follow_the_reference():
    # CHAT_SAMPLER::InitPrefillTemplate, reference src/Utils/CLI_params.cpp:1999-2005 (the four printf templates)
    assert kf.chatml_prompt("hi", enable_thinking=True) == "<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n"
    assert kf.chatml_prompt("hi", "be brief", enable_thinking=True) == "<|im_start|>system\nbe brief<|im_end|>\n<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n"
    assert kf.chatml_prompt("hi") == "<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n<think>\n\n</think>\n\n"
    assert kf.chatml_prompt("hi", "be brief") == ("<|im_start|>system\nbe brief<|im_end|>\n<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n"
                                                   "<think>\n\n</think>\n\n")
    # CHAT_SAMPLER::toChatML, :2010-2031; the example in the reference's own comment (src/TokenSet/TokenSet.cpp:809)
    lines = [("system", "You are a dog."), ("user", "Hello"), ("assistant", "Fine")]
    assert kf.chatml_render(lines) == ("<|im_start|>system\nYou are a dog.<|im_end|>\n<|im_start|>user\nHello<|im_end|>\n<|im_start|>assistant\n"
                                       "<think>\n\n</think>\n\nFine<|im_end|>\n")
    assert kf.chatml_render(lines, enable_thinking=True).endswith("<|im_start|>assistant\n\n\nFine<|im_end|>\n")
    assert kf.chatml_render([]) == ""


def test_chatml_equals_the_reference_functions_compiled_from_its_tree():
    # CHAT_SAMPLER::InitPrefillTemplate / toChatML (reference src/Utils/CLI_params.cpp:1990-2031) from oracle/_ref/libkoifish_refcpu.so
    import oracle_lib as ol
    if ol.refcpu() is None:
        pytest.skip("oracle/_ref/libkoifish_refcpu.so not built (reference tree absent at build time)")
    for thinking in (False, True):
        user_t, sys_t = ol.refcpu_prefill_templates(thinking)
        for user, system in (("hi", None), ("What is 2 + 2?", "You are a helpful assistant."), ("多行\n文本 100%", "sys % s")):
            want = (sys_t.replace("%s", "{}").format(system, user)) if system else user_t.replace("%s", "{}").format(user)
            assert kf.chatml_prompt(user, system, enable_thinking=thinking) == want
        for lines in ([], [("user", "Hello")], [("system", "You are a dog."), ("user", "Hello"), ("assistant", "Fine")],
                      [("user", "a"), ("assistant", "b"), ("user", "c"), ("assistant", "d\n")]):
            assert kf.chatml_render(lines, enable_thinking=thinking) == ol.refcpu_tochatml(lines, thinking)


def test_chat_prompt_round_trip(tok, gold):
    p = kf.chatml_prompt("What is the capital of France?", "You are a helpful assistant.")
    ids = tok.encode(p)
    sp = gold["specials"]
    assert ids[0] == sp["<|im_start|>"] and ids.count(sp["<|im_start|>"]) == 3 and ids.count(sp["<|im_end|>"]) == 2
    assert sp["<think>"] in ids and sp["</think>"] in ids
    assert tok.decode(ids) == p
    assert tok.decode(ids, skip_special_tokens=True) == p.replace("<|im_start|>", "").replace("<|im_end|>", "")


def test_random_strings_against_the_library_live(tok, tok_l3):
    tokenizers = pytest.importorskip("tokenizers")
    hf = tokenizers.Tokenizer.from_file(os.path.join(GOLD, "tokenizer.json"))
    hf_l3 = tokenizers.Tokenizer.from_file(os.path.join(GOLD, "llama3style", "tokenizer.json"))
    rng = random.Random(2026)
    alphabet = (list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ") * 2 + list("0123456789") * 2 + list("    \t\n\n\r") * 3 +
                list("'''.,;:!?-_()[]{}<>|/\\\"@#$%^&*+=~`") + list("éèüñß̧́̈абв中文天あア가각"
                                                               "الَक़ก้ 　 ​²½①\U0001f600\U0001f389ſK") +
                ["<|im_start|>", "<|im_end|>", "<|endoftext|>", "<think>", "</think>", "'s", "'T", "'re", "'LL", " the", " and", "ing"])
    for k in range(1500):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40)))
        want = hf.encode(s, add_special_tokens=False).ids
        got = tok.encode(s)
        assert got == want, repr(s)
        assert tok_l3.encode(s) == hf_l3.encode(s, add_special_tokens=False).ids, repr(s)
        assert tok.decode(got) == hf.decode(want, skip_special_tokens=False), repr(s)
        if k % 10 == 0:
            cut = want[:rng.randint(0, len(want))]
            assert tok.decode(cut, skip_special_tokens=True) == hf.decode(cut, skip_special_tokens=True), repr(s)


def test_streaming_decode_yields_whole_characters_and_sums_to_decode(tok, gold):
    rng = random.Random(3)
    texts = [c["text"] for c in gold["cases"]] + ["天\U0001f600é" * 3]
    for text in texts:
        ids = tok.encode(text)
        pieces = list(tok.stream(ids))
        assert len(pieces) == len(ids) + 1
        assert "".join(pieces) == tok.decode(ids), repr(text)
        for p in pieces[:-1]:
            assert "�" not in p or "�" in text, repr(text)  # no broken characters while the sequence is a real encoding
    # arbitrary id sequences (cut characters, stray continuation bytes): still the same text as decode() in one go, pieces still valid UTF-8
    usable = [i for i in range(tok.vocab_size) if "\u0100" not in tok.id_to_token(i)]  # U+0100 is byte 0: NUL would end the returned C string
    for _ in range(300):
        ids = [rng.choice(usable) for _ in range(rng.randint(0, 30))]
        for skip in (False, True):
            pieces = list(tok.stream(ids, skip_special_tokens=skip))
            assert "".join(pieces) == tok.decode(ids, skip_special_tokens=skip), ids


def test_against_the_reference_tokenizer_compiled_from_its_tree(tok, gold):
    # the reference's own src/TokenSet/HF_Tokenizer.cpp (+ Dictionary.cpp and the oniguruma / utf8proc it vendors) compiled where it lies
    # (oracle/_ref/libkoifish_reftok.so, oracle/ref_tokenizer.cpp) on the same tokenizer.json.  The reference does not build the NFC normalizer this
    # file declares (NFKC is its only normalisation form, HF_Tokenizer.cpp:202-215) -- it differs from the HF library on exactly the 14 golden texts
    # that NFC changes -- so it is fed the normalised text; everything after normalisation (added tokens, the Split pattern through Oniguruma,
    # byte-level BPE, decoding) must agree id for id and byte for byte.
    import oracle_lib as ol
    ref = ol.reftok(open(os.path.join(GOLD, "tokenizer.json")).read())
    if ref is None:
        pytest.skip("oracle/_ref/libkoifish_reftok.so not built (reference tree absent at build time)")
    unnormalised = 0
    for c in gold["cases"]:
        text = c["text"]
        ids = tok.encode(text)
        assert ref.encode(kf.nfc(text)) == ids, repr(text)
        unnormalised += ref.encode(text) != ids
        assert ref.decode(ids) == tok.decode(ids), repr(text)
        assert ref.decode(ids, True) == tok.decode(ids, skip_special_tokens=True), repr(text)
    assert unnormalised == 14
    rng = random.Random(77)
    alphabet = (list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ") * 2 + list("0123456789") * 2 + list("    \t\n\n\r") * 3 +
                list("'''.,;:!?-_()[]{}<>|/\\\"@#$%^&*+=~`") + list("éèüñßабв中文天あア가 　²½①\U0001f600\U0001f389") +
                ["<|im_start|>", "<|im_end|>", "<|endoftext|>", "<think>", "</think>", "'s", "'T", "'re", "'LL", " the", " and", "ing"])
    for _ in range(800):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40)))
        ids = tok.encode(s)
        assert ref.encode(kf.nfc(s)) == ids, repr(s)
        assert ref.decode(ids) == tok.decode(ids), repr(s)
