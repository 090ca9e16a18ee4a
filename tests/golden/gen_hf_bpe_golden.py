"""Builds a Llama-3-style HuggingFace tokenizer.json offline (byte-level BPE, Split pre-tokenizer with the cl100k / Llama-3 regex,
ignore_merges = true, added special tokens) with the Python `tokenizers` package — the bindings of the same Rust crate the reference
calls (`tokenizers = "0.23.1"`, crates/tokenizer/src/huggingface.rs:310-316; here 0.22.x) — and records its encodings as golden vectors.
Two variants: ignore_merges true (Llama 3) and false (plain HF BPE).  Run from the repo root: python tests/golden/gen_hf_bpe_golden.py"""
import json
import os
import random

from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
SPECIALS = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]
WORDS = ("the quick brown fox jumps over lazy dog cache aware router prefix radix tree worker tenant load balance token hello world "
         "naïve café crème brûlée 你好 世界 调度 été ça Straße größer 12345 2024 3.14159 don't I'll we've they're it's x=y+1 foo_bar() "
         "{\"key\": [1, 2, 3]} https://example.com/a?b=c&d=e \t tab\n newline\r\n crlf   spaces   👋🌍 emoji ＡＢＣ ١٢٣").split(" ")


def corpus(n, seed):
    r = random.Random(seed)
    out = []
    for _ in range(n):
        k = r.randrange(1, 40)
        out.append(" ".join(r.choice(WORDS) for _ in range(k)))
    return out


def build(ignore_merges):
    tok = Tokenizer(models.BPE(ignore_merges=ignore_merges))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PATTERN), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=2500, special_tokens=[], initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(corpus(4000, 7), trainer)
    tok.add_special_tokens(SPECIALS)
    return tok


def main():
    texts = corpus(150, 11) + ["", " ", "a", "  leading", "trailing  ", "\n\n\n", " \n \n x", "tabs\t\tand  \n  mixed \r\n", "'s 'T 'Re 'VE 'm 'LL 'd 'x",
                               "1 12 123 1234 12345 1234567", "<|begin_of_text|>hello<|eot_id|>", "a<|start_header_id|>user<|end_header_id|>\n\nhi<|eot_id|>",
                               "<|begin_of_text|<|eot_id|>>", "x" * 300, "ab" * 200, "你好" * 50, "👋" * 20, "mixed 你好world123 ünï"]
    out = {}
    for name, im in (("hf_llama3_style_tokenizer.json", True), ("hf_plain_bpe_tokenizer.json", False)):
        tok = build(im)
        path = os.path.join(HERE, name)
        tok.save(path)
        enc = tok.encode_batch(texts, add_special_tokens=False)
        out[name] = [e.ids for e in enc]
    json.dump({"texts": texts, "ids": out}, open(os.path.join(HERE, "hf_bpe_vectors.json"), "w"))
    print({k: sum(len(x) for x in v) for k, v in out.items()})


if __name__ == "__main__":
    main()
