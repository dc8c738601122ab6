"""Golden vectors for csrc/TokenSet, produced with the HF `tokenizers` library -- the crate the reference's
src/TokenSet/HF_Tokenizer.cpp ports.  No Qwen3 tokenizer.json exists offline, so this script TRAINS a small byte-level BPE with exactly the
pipeline a Qwen3 tokenizer.json declares (NFC; Split(<pattern>, Isolated) + ByteLevel(use_regex = false); BPE; ByteLevel decoder; the ChatML
added tokens) and writes
    tests/golden/tokenizer/tokenizer.json, tokenizer_config.json      -- the Qwen-style fixture (pattern with \\p{N})
    tests/golden/tokenizer/llama3style/tokenizer.json                  -- the same with \\p{N}{1,3} and ignore_merges
    tests/golden/tokenizer/cases.json                                  -- text -> pieces / ids / decoded text, as the library computes them
Run here (needs `tokenizers`):  python tests/golden/make_tokenizer_golden.py"""
import json
import os
import random

from tokenizers import AddedToken, Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, processors, trainers

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tokenizer")
PAT_QWEN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
PAT_LLAMA3 = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"

CORPUS = [
    "The quick brown fox jumps over the lazy dog. It's a test, isn't it? We'll see; they've said I'm right and you'd agree.",
    "Koifish is a quantized inference framework: 4-bit, 2-bit and 1-bit weights with group size 128 on 1 GPU or 8 GPUs.",
    "def forward(self, x):\n    return self.proj(torch.relu(x)) + 1.0e-6  # residual\n\n\nclass Model(nn.Module):\n\tpass\n",
    "天命之谓性，率性之谓道，修道之谓教。今天天气很好，我们去公园散步吧！2024年10月18日。",
    "こんにちは世界。これはテストです。カタカナとひらがな。",
    "안녕하세요 세계. 이것은 시험입니다.",
    "Привет, мир! Это тест токенизатора. Ёлка, съешь ещё этих мягких булок.",
    "Ça va très bien, merci. Où est la bibliothèque? L'été à Zürich: Äpfel, Öl, Übung, straße, naïve café.",
    "مرحبا بالعالم. هذا اختبار. مَرْحَبًا",
    "नमस्ते दुनिया। यह एक परीक्षण है।",
    "\U0001f600 emoji \U0001f389\U0001f389 and symbols ∑∫√ ≠ ≤ → ← ©®™ §¶ … — – ‘quotes’ “double”",
    "1234567890 3.14159 2,718,281 0x1F 1e-5 100% $42.00 #hashtag @user a_b-c/d\\e",
    "<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n<|im_start|>user\nWhat is the capital of France?<|im_end|>\n<|im_start|>assistant\n",
    "    indented     text\twith\ttabs  \n  and trailing spaces   \n\n\nnew paragraph\r\nwindows line\r\n\r\n",
    "HELLO WORLD I'M SHOUTING AND YOU'RE NOT; HE'S, SHE'LL, THEY'VE, WE'D, DON'T",
]

CASES = [
    "", " ", "  ", "\n", "a", "Hello", "Hello world", " Hello  world ", "Hello, world!", "It's they're we've I'm you'll he'd don't",
    "IT'S THEY'RE WE'VE I'M YOU'LL HE'D DON'T 'Twas 'sup 'tis", "'s't're've'm'll'd", "x'S y'T z'RE q'Ve w'M e'lL r'D", "'ſ long s", "a'ſb", "'K kelvin",
    "12345", "1 22 333 4444 55555 666666 7777777", "3.14159 and 2,718", "abc123def456", "٣٤٥ ١٢ Ⅻ ² ½ ①②③",
    "price: $42.00 (incl. 7% VAT)!!!",
    "a  b   c    d", "tab\tseparated\t\tvalues", "line1\nline2\n\nline4", "x \n y", "x\n \ny", "a \r\n\r\n b", "trailing   ", "   leading", " \n ", "\n\n\n",
    "  \n  \n  x", "a\t \n\t b", "end.\n", "end. \n\n", "!!! ??? ... ---", " !!!", "  !!!", "a !b", "a  !b", "(parens) [brackets] {braces} <angle>",
    "semi;colon:quote\"apos'", "... \n\nnext", "foo();\n\n\nbar();", "if (x) {\n    y();\n}\n", "#include <stdio.h>\nint main() { return 0; }",
    "天命之谓性", "今天 天气 很好", "中文English混合text", "日本語のテキストです。",
    "한국어 텍스트", "Привет мир", "Ελληνικά", "עברית שלום",
    "مرحبا بالعالم", "مَرْحَبًا", "नमस्ते दुनिया", "ไทย ภาษา ก้ำ",
    "café naïve Zürich", "café naïve Zürich", "Å Å Å", "각 한 각", "ộ ộ ộ",
    "é́", "́abc", "à́̂̃", "q̣̇ q̣̇", "̈́ ̀ ́ ̓ ʹ ; ·", "क़ ড় ଡ଼ གྷ ⫝̸ יִ שׁ", "\U0001d15e \U0001d1bb \U0002f800",
    "ୋ ୈ ේා ဦ ো ொ", "ﬁ ligature ﬂ", "①②③ ㈱ ㍿", "\U0001f600\U0001f603\U0001f604",
    "\U0001f468‍\U0001f469‍\U0001f467‍\U0001f466 family", "\U0001f1eb\U0001f1f7 flag", "a\U0001f600b", "emoji \U0001f389 end",
    "nbsp here", "ideographic　space", "line sep", "para sep", "nelchar", "fschar", "uschar", "zwsp​char", "thin space", "a  b",
    "　　x", "bom﻿char", "mongolian᠎vowel", "ogham space", "<|im_start|>", "<|im_start|>user\nhi<|im_end|>\n", "<|endoftext|>", "a<|im_end|>b",
    "<|im_end|><|im_end|>", "<|im_end", "<think>\n\n</think>\n\n",
    "<|im_start|>assistant\n<think>\nLet me think.\n</think>\n\nParis.<|im_end|>", "< |im_start|>", "<<|im_start|>>", "<think><think>", "<tool_call>{\"name\": \"f\"}</tool_call>",
    "The quick brown fox jumps over the lazy dog.", "Koifish 4-bit quantized inference on B200", "https://example.com/path?q=1&r=2#frag", "user@example.com",
    "snake_case camelCase PascalCase kebab-case", "C++ C# F# .NET node.js", "1st 2nd 3rd 4th", "x² + y² = z²", "α + β = γ", "10km/h 5°C 3µs",
    "a" * 40, " " * 33, "ab " * 20, "\t\t\t", "\r", "\r\n", "a\rb", "mixed \t \n \r\n ws", "ÀÉÎÕÜ àéîõü", "ßẞ ſ",
    "İstanbul ılık", "ǆ ǅ Ǆ", "Ω Ω K K", "́ͅ ᾴ ᾴ",
]


def build(pattern, ignore_merges, vocab_size, seed_texts):
    tok = Tokenizer(models.BPE(ignore_merges=ignore_merges))
    tok.normalizer = normalizers.NFC()
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(pattern), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    tok.post_processor = processors.ByteLevel(trim_offsets=False)
    trainer = trainers.BpeTrainer(vocab_size=vocab_size, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False, special_tokens=[])
    tok.train_from_iterator(seed_texts, trainer)
    tok.add_special_tokens([AddedToken(s, special=True, normalized=False) for s in ("<|endoftext|>", "<|im_start|>", "<|im_end|>")])
    tok.add_tokens([AddedToken(s, special=False, normalized=False) for s in ("<think>", "</think>", "<tool_call>", "</tool_call>")])
    return tok


def main():
    os.makedirs(os.path.join(HERE, "llama3style"), exist_ok=True)
    rng = random.Random(42)
    texts = list(CORPUS)
    for _ in range(6):  # repetition gives the trainer merges to find
        texts += [" ".join(rng.sample(t.split(" "), k=max(1, len(t.split(" ")) // 2))) for t in CORPUS]
    qwen = build(PAT_QWEN, False, 1400, texts)
    qwen.save(os.path.join(HERE, "tokenizer.json"), pretty=False)
    with open(os.path.join(HERE, "tokenizer_config.json"), "w") as f:
        json.dump({"eos_token": "<|im_end|>", "pad_token": "<|endoftext|>", "bos_token": None, "tokenizer_class": "Qwen2Tokenizer"}, f)
    llama = build(PAT_LLAMA3, True, 900, texts)
    llama.save(os.path.join(HERE, "llama3style", "tokenizer.json"), pretty=False)
    out = {"tokenizers_version": __import__("tokenizers").__version__, "cases": []}
    split_q = pre_tokenizers.Split(Regex(PAT_QWEN), behavior="isolated")
    split_l = pre_tokenizers.Split(Regex(PAT_LLAMA3), behavior="isolated")
    nfc = normalizers.NFC()
    for text in CASES + CORPUS:
        enc, encl = qwen.encode(text, add_special_tokens=False), llama.encode(text, add_special_tokens=False)
        out["cases"].append({
            "text": text,
            "nfc": nfc.normalize_str(text),
            "pieces": [p for p, _ in split_q.pre_tokenize_str(nfc.normalize_str(text))],
            "pieces_llama3": [p for p, _ in split_l.pre_tokenize_str(nfc.normalize_str(text))],
            "ids": enc.ids, "ids_llama3": encl.ids,
            "decoded": qwen.decode(enc.ids, skip_special_tokens=False),
            "decoded_skip_special": qwen.decode(enc.ids, skip_special_tokens=True),
        })
    # decoding id sequences that cut a multi-byte character: from_utf8_lossy behaviour
    ids = qwen.encode("天\U0001f600é", add_special_tokens=False).ids
    out["lossy"] = [{"ids": ids[a:b], "decoded": qwen.decode(ids[a:b], skip_special_tokens=False)} for a in range(len(ids)) for b in range(a + 1, len(ids) + 1)]
    out["vocab_size"] = qwen.get_vocab_size(with_added_tokens=True)
    out["specials"] = {s: qwen.token_to_id(s) for s in ("<|endoftext|>", "<|im_start|>", "<|im_end|>", "<think>", "</think>")}
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(out, f, ensure_ascii=True)
    print("wrote", HERE, "cases", len(out["cases"]), "vocab", out["vocab_size"])


if __name__ == "__main__":
    main()
