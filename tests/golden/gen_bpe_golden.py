"""Generates the BPE golden fixtures.  Run in the build container only:  python tests/golden/gen_bpe_golden.py

Source of truth: Python `tiktoken` 0.12.0 (OpenAI's Rust CoreBPE) with the reference's cl100k pattern
(crates/tokenizer/src/tiktoken.rs:28) over a SYNTHETIC vocabulary trained offline -- no real vocab file exists on
this machine.  The reference reaches the same algorithm through the un-vendored crate tiktoken-rs 0.9.1
(crates/tokenizer/Cargo.toml:39; call site tiktoken.rs:460 `encode_with_special_tokens`).
Outputs (committed):  synth_vocab.tiktoken  (base64(token) rank lines, the format load_tiktoken_bpe parses, tiktoken.rs:346-367)
                      bpe_vectors.json      (texts + expected token ids, incl. special tokens)
"""
import base64
import glob
import json
import os
import random

import tiktoken
from tiktoken import _educational as ed

HERE = os.path.dirname(os.path.abspath(__file__))
PAT = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
SPECIALS = {"<|endoftext|>": 3000, "<|im_start|>": 3001, "<|im_end|>": 3002, "<|fim_prefix|>": 3003, "<|im_start|>x": 3004}

UNICODE_STRESS = [
    "你好世界，今天天气怎么样？我们去公园散步吧。",
    "こんにちは世界。トークナイザーのテストです。",
    "안녕하세요 세계, 토크나이저 테스트입니다.",
    "Привет, мир! Это тест токенизатора.",
    "مرحبا بالعالم، هذا اختبار.",
    "naïve café déjà vu — coöperate façade",
    "Ünïcödé ſtrange 'ſ and K (kelvin) 'S 'LL 'Re",
    "emoji \U0001f44b\U0001f30d\U0001f680 mixed \U0001f469‍\U0001f469‍\U0001f467‍\U0001f466 zwj",
    "math ∑∫√≠≤ ½ ⅓ ² ³ ① Ⅷ ٣٤٥ १२३",
    "tabs\tand nbsp emspace　ideographic", "line\r\nbreaks\n\n\nmany\r\r end   ",
    "   leading spaces", "trailing spaces    ", "a  b   c    d", "x\n y\n  z\n\t\tw", "1234567890 12 345 6789 3.14159 1,000,000",
    "don't I'll we've they're it's I'M HE'D 'tis 'twas",
    "snake_case camelCase PascalCase kebab-case SCREAMING_CASE", "http://worker-12:8000/v1/chat/completions?x=1&y=2#frag",
    "{\"model\": \"m\", \"messages\": [{\"role\": \"user\", \"content\": \"hi\"}]}",
    "!!!???...,,,;;;:::---___***///\\\\\\", "     \n     \n", "\n", " ", "", "a", "'", "''s", "'s's", "é", "ñandú",
    "\U0001d518\U0001d52b\U0001d526\U0001d520\U0001d52c\U0001d521\U0001d522 \U0001d544\U0001d538\U0001d54bℍ", "nextline para ",
]


def corpus():
    rng = random.Random(7)
    files = sorted(glob.glob("/usr/lib/python3*/[a-m]*.py"))[:60]
    chunks = []
    for f in files:
        try:
            chunks.append(open(f, encoding="utf-8", errors="ignore").read()[:6000])
        except OSError:
            pass
    text = "\n".join(chunks)
    words = ["the quick brown fox jumps over the lazy dog", "You are a helpful assistant.", "Summarize the following document in three bullet points:",
             "Translate to French:", "What is the capital of", "Write a Python function that", "Explain quantum computing to a five year old."]
    extra = "\n".join(rng.choice(words) + " " + " ".join(rng.choice(words).split()[: rng.randint(1, 6)]) + f" (variant {i})" for i in range(600))
    return text[:220000] + "\n" + extra + "\n" + "\n".join(UNICODE_STRESS * 6)


def main():
    data = corpus()
    ranks = ed.bpe_train(data, 3000, PAT, visualise=None)
    assert len(ranks) == 3000
    with open(os.path.join(HERE, "synth_vocab.tiktoken"), "w") as f:
        for tok, r in sorted(ranks.items(), key=lambda kv: kv[1]):
            f.write(base64.b64encode(tok).decode() + " " + str(r) + "\n")
    enc = tiktoken.Encoding("synth", pat_str=PAT, mergeable_ranks=ranks, special_tokens=SPECIALS)
    rng = random.Random(11)
    lines = [ln for ln in data.split("\n") if ln.strip()]
    texts = list(UNICODE_STRESS)
    texts += rng.sample(lines, 150)
    texts += ["\n".join(rng.sample(lines, 8)) for _ in range(30)]
    texts += ["<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n<|im_start|>user\n" + rng.choice(lines) + "<|im_end|>\n<|im_start|>assistant\n"
              for _ in range(20)]
    texts += ["<|endoftext|>", "<|endoftext|><|endoftext|>", "a<|im_end|>b", "<|im_start|>xyz", "<|im_star", "<|unknown|> <|im_end|", "<<|im_end|>>"]
    texts += [" " * n for n in (2, 3, 7, 33, 100)] + ["\n" * 40, "ab" * 300, "!" * 90, "0" * 50, "z" * 200, " x" * 120, "é" * 70]
    alphabet = "abc 12\n\t'.,-_äß你\U0001f44b"
    texts += ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 120))) for _ in range(120)]
    cases = [{"text": t, "ids": enc.encode(t, allowed_special="all")} for t in texts]
    json.dump({"generator": "tiktoken " + tiktoken.__version__, "pattern": PAT, "specials": SPECIALS, "cases": cases},
              open(os.path.join(HERE, "bpe_vectors.json"), "w"), ensure_ascii=True)
    print("vocab", len(ranks), "cases", len(cases), "tokens", sum(len(c["ids"]) for c in cases))


if __name__ == "__main__":
    main()
