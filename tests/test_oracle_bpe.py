"""BPE oracle (oracle/bpe_ref.py) pinned against the golden vectors produced by Python tiktoken (OpenAI's Rust CoreBPE)
and, when importable, against live tiktoken on random text.  CPU only."""
import json
import os
import random

import pytest

from oracle import bpe_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load():
    g = json.load(open(os.path.join(GOLD, "bpe_vectors.json")))
    ranks = bpe_ref.load_tiktoken_bpe(os.path.join(GOLD, "synth_vocab.tiktoken"))
    return g, ranks


def test_pattern_is_the_reference_pattern():
    g, _ = _load()
    assert g["pattern"] == bpe_ref.CL100K_BASE_PATTERN   # crates/tokenizer/src/tiktoken.rs:28


def test_oracle_matches_golden_vectors():
    g, ranks = _load()
    enc = bpe_ref.CoreBPE(ranks, g["specials"])
    assert len(g["cases"]) > 300
    for c in g["cases"]:
        assert enc.encode_with_special_tokens(c["text"]) == c["ids"], c["text"][:80]


def test_oracle_matches_live_tiktoken_on_random_text():
    tiktoken = pytest.importorskip("tiktoken")
    g, ranks = _load()
    live = tiktoken.Encoding("synth", pat_str=g["pattern"], mergeable_ranks=ranks, special_tokens=g["specials"])
    enc = bpe_ref.CoreBPE(ranks, g["specials"])
    rng = random.Random(3)
    alphabet = "abcdefghij XYZ 0123\n\t\r'.,!?-_()<|>im_startend" + "éß你好\U0001f44b 　"
    for _ in range(400):
        t = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 160)))
        assert enc.encode_with_special_tokens(t) == live.encode(t, allowed_special="all"), repr(t)
