"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Restatement of the tokenize step of the hot path for tiktoken-style models:
    Tokenizer::encode → TiktokenTokenizer::encode → CoreBPE::encode_with_special_tokens
    (crates/tokenizer/src/lib.rs:85, tiktoken.rs:444-462, pattern CL100K_BASE_PATTERN tiktoken.rs:28,
     vocab loader load_tiktoken_bpe tiktoken.rs:346-367).
The BPE arithmetic lives in the un-vendored crate tiktoken-rs 0.9.1 (crates/tokenizer/Cargo.toml:39), a port of OpenAI's
tiktoken; this file restates tiktoken's published algorithm:
  1. split the text at special-token strings (leftmost match; at one position the LONGEST special wins here — the crate's
     alternation order comes from HashMap iteration and is arbitrary);
  2. split ordinary text with the regex (leftmost-first alternation, backtracking look-ahead `\\s+(?!\\S)`);
  3. a piece that is a vocab entry is one token; otherwise byte-pair merge: repeatedly merge the LEFTMOST adjacent pair
     whose concatenated bytes have the lowest rank (`_byte_pair_merge`).
Pinned by tests/golden/bpe_vectors.json (generated with Python tiktoken 0.12.0, OpenAI's own Rust core, over the synthetic
vocabulary tests/golden/synth_vocab.tiktoken — no real vocabulary file exists offline) and, when the package
is importable, against live tiktoken on random text.
"""
import base64

import regex

CL100K_BASE_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")

_RANK_MAX = 0xFFFFFFFF


def load_tiktoken_bpe(path):
    """tiktoken.rs:346-367 — lines of `base64(token) rank`."""
    ranks = {}
    for line in open(path).read().splitlines():
        if not line:
            continue
        tok, rank = line.split()
        ranks[base64.b64decode(tok)] = int(rank)
    return ranks


def byte_pair_merge(ranks, piece: bytes):
    """tiktoken `_byte_pair_merge`: parts = [(start, rank of the pair starting there)]."""
    parts = []
    min_rank = (_RANK_MAX, None)
    for i in range(len(piece) - 1):
        r = ranks.get(piece[i:i + 2], _RANK_MAX)
        if r < min_rank[0]:
            min_rank = (r, i)
        parts.append([i, r])
    parts.append([len(piece) - 1, _RANK_MAX])
    parts.append([len(piece), _RANK_MAX])

    def get_rank(i):
        if i + 3 < len(parts):
            return ranks.get(piece[parts[i][0]:parts[i + 3][0]], _RANK_MAX)
        return _RANK_MAX

    while min_rank[0] != _RANK_MAX:
        i = min_rank[1]
        if i > 0:
            parts[i - 1][1] = get_rank(i - 1)
        parts[i][1] = get_rank(i)
        del parts[i + 1]
        min_rank = (_RANK_MAX, None)
        for j in range(len(parts) - 1):
            if parts[j][1] < min_rank[0]:
                min_rank = (parts[j][1], j)
    return parts


def byte_pair_encode(ranks, piece: bytes):
    if len(piece) == 1:
        return [ranks[piece]]
    parts = byte_pair_merge(ranks, piece)
    return [ranks[piece[parts[i][0]:parts[i + 1][0]]] for i in range(len(parts) - 1)]


class CoreBPE:
    def __init__(self, ranks, special_tokens, pattern=CL100K_BASE_PATTERN):
        self.ranks, self.special, self.re = ranks, dict(special_tokens), regex.compile(pattern)

    def pieces(self, text):
        return [m.group(0) for m in self.re.finditer(text)]

    def encode_ordinary(self, text):
        out = []
        for piece in self.pieces(text):
            b = piece.encode("utf-8")
            if b in self.ranks:
                out.append(self.ranks[b])
            else:
                out.extend(byte_pair_encode(self.ranks, b))
        return out

    def _next_special(self, text, start):
        best = None
        for s in self.special:
            k = text.find(s, start)
            if k >= 0 and (best is None or k < best[0] or (k == best[0] and len(s) > len(best[1]))):
                best = (k, s)
        return best

    def encode_with_special_tokens(self, text):
        out, start = [], 0
        while True:
            nxt = self._next_special(text, start)
            end = nxt[0] if nxt else len(text)
            out.extend(self.encode_ordinary(text[start:end]))
            if not nxt:
                return out
            out.append(self.special[nxt[1]])
            start = nxt[0] + len(nxt[1])
