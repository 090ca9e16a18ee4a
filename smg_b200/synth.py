"""Synthetic fleets, KV-index populations and request batches shaped like BASELINE.json's configs
(SURVEY.md §8d).  Pure numpy data generation shared by bench.py and the parity tests; no routing logic.

Token ids ~ U[0, 50000) like the reference's own bench generator (model_gateway/benches/radix_tree_benchmark.rs:82-87).
"""
import numpy as np

VOCAB = 50000
PAGE = 16


def worker_urls(n):
    return [f"http://worker-{i}:8000" for i in range(n)]


def gen_sequences(n_seq, n_tokens, seed, vocab=VOCAB):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, vocab, size=(n_seq, n_tokens), dtype=np.uint32)


def poisson_loads(n_workers, lam, seed, clip=64):
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    return np.minimum(rng.poisson(lam, size=n_workers), clip).astype(np.uint64)


def gen_queries(seqs, batch, seed, full_hit=0.8, partial=0.1, vocab=VOCAB, block=PAGE):
    """Config-2 mix: `full_hit` stored full paths, `partial` stored prefix of block·U[1, P-1] tokens then novel tokens,
    the rest novel.  Returns uint32 [batch, T]."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    n_seq, T = seqs.shape
    P = T // block
    kind = rng.random(batch)
    pick = rng.integers(0, n_seq, size=batch)
    out = rng.integers(0, vocab, size=(batch, T), dtype=np.uint32)       # novel by default
    full = kind < full_hit
    out[full] = seqs[pick[full]]
    part = (kind >= full_hit) & (kind < full_hit + partial)
    if P > 1:
        keep = rng.integers(1, P, size=batch) * block
        for i in np.nonzero(part)[0]:
            out[i, :keep[i]] = seqs[pick[i], :keep[i]]
    return out


def zipf_prefix_queries(n_prompts, batch, seed, s=1.1, vocab=VOCAB, block=PAGE, min_pages=8, max_pages=96, fresh_pages=(1, 16)):
    """Config-3 shape: `n_prompts` system prompts of block·U[min,max] tokens with Zipf(s) popularity, each request =
    a system prompt + block·U[1,16] fresh tokens.  Returns (prompts list, ragged tokens, offsets, prompt ids)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(min_pages, max_pages + 1, size=n_prompts) * block
    prompts = [rng.integers(0, vocab, size=int(n), dtype=np.uint32) for n in lens]
    w = 1.0 / np.arange(1, n_prompts + 1) ** s
    w /= w.sum()
    ids = rng.choice(n_prompts, size=batch, p=w)
    fresh = rng.integers(fresh_pages[0], fresh_pages[1] + 1, size=batch) * block
    reqs = [np.concatenate([prompts[i], rng.integers(0, vocab, size=int(f), dtype=np.uint32)]) for i, f in zip(ids, fresh)]
    offsets = np.zeros(batch + 1, dtype=np.uint32)
    np.cumsum([len(r) for r in reqs], out=offsets[1:])
    return prompts, np.concatenate(reqs).astype(np.uint32), offsets, ids


def ragged(batch_2d):
    """[B, T] → (flat tokens, offsets u32[B+1])."""
    B, T = batch_2d.shape
    return np.ascontiguousarray(batch_2d.reshape(-1)), (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
