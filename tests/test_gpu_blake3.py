"""GPU parity for the mesh path hashes (a16): batched BLAKE3 on the device against the official vectors
(tests/golden/blake3_vectors.json), the oracle on random ragged batches, and the hash_index side effect of the tree-mode
select calls (cache_aware.rs:397-401, 420-424, 881-886, 950-956; test_apply_repair_page_seeds_hash_index :1244-1362)."""
import json
import os

import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth

pytestmark = pytest.mark.gpu
V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "blake3_vectors.json")))
CFG = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)


def _pol(**kw):
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    return CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), **kw)


def test_official_vectors_as_one_ragged_batch():
    pol = _pol()
    texts = [bytes(i % 251 for i in range(v["len"])) for v in V["bytes"]]
    # hash_node_paths takes str; feed the raw bytes through the C ABI directly (any byte string is hashable)
    import ctypes as C
    offs = np.zeros(len(texts) + 1, np.uint32)
    np.cumsum([len(t) for t in texts], out=offs[1:])
    blob = np.frombuffer(b"".join(texts), dtype=np.uint8).copy()
    out = np.zeros(len(texts), np.uint64)
    pol._h.call("smgx_hash_node_paths", blob.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), len(texts), out.ctypes.data_as(C.c_void_p))
    assert [int(x) for x in out] == [int(v["path_hash"]) for v in V["bytes"]]


def test_reference_call_shapes():
    pol = _pol()
    got = pol.hash_node_paths([v["text"] for v in V["node_paths"]])
    assert [int(x) for x in got] == [int(v["path_hash"]) for v in V["node_paths"]]
    toks = []
    for v in V["token_paths"]:
        if v["tokens"] is not None:
            toks.append(v["tokens"])
        elif v["gen"][2] == "alt":
            toks.append([0, 0xFFFFFFFF] * (v["gen"][1] // 2))
        else:
            toks.append(list(range(v["gen"][0], v["gen"][0] + v["gen"][1])))
    got = pol.hash_token_paths(toks)
    assert [int(x) for x in got] == [int(v["path_hash"]) for v in V["token_paths"]]
    assert all(int(x) != 0 for x in got)


@pytest.mark.parametrize("seed", [1, 2])
def test_random_ragged_batches_match_oracle(seed):
    pol = _pol(max_batch=8192)
    rng = np.random.default_rng(seed)
    reqs = [rng.integers(0, 2**32, size=int(n), dtype=np.uint64).astype(np.uint32) for n in rng.integers(0, 3000, size=700)]
    reqs += [rng.integers(0, 50000, size=512, dtype=np.uint32) for _ in range(300)]          # the config-2 shape
    got = pol.hash_token_paths(reqs)
    for r, g in zip(reqs, got):
        assert int(g) == orc.hash_token_path(r)
    texts = ["".join(chr(int(c)) for c in rng.choice([0x61, 0x20, 0xE9, 0x4F60, 0x1F44B], size=int(n))) for n in rng.integers(0, 1500, size=200)]
    got = pol.hash_node_paths(texts)
    for t, g in zip(texts, got):
        assert int(g) == orc.hash_node_path(t)


def test_hash_index_side_effect_matches_oracle():
    from smg_b200 import BasicWorker
    urls = synth.worker_urls(4)
    pol = _pol()
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    rng = np.random.default_rng(5)
    trunk = rng.integers(0, 50000, size=64, dtype=np.uint32)
    reqs = [np.concatenate([trunk[: 16 * int(rng.integers(0, 5))], rng.integers(0, 50000, size=16 * int(rng.integers(1, 4)), dtype=np.uint32)]) for _ in range(40)]
    reqs.append(np.arange(32, dtype=np.uint32))                     # the reference test's request (:1347)
    pol.select_worker_batch(ws, reqs)
    for r in reqs:
        op.select_worker(tokens=r)
    want = op.hash_index("tokens")
    assert pol.hash_index_size("tokens") == len(want)
    for h, prefix in want.items():
        assert pol.hash_index_get(h, "tokens") == prefix
    texts = ["the quick brown fox jumps over the lazy dog", "the quick brown cat", "the quick", "你好世界 and more", "你好吗"]
    pol.select_worker_batch_request_text(ws, texts)
    for t in texts:
        op.select_worker(request_text=t)
    want = op.hash_index("text")
    assert pol.hash_index_size("text") == len(want)
    for h, prefix in want.items():
        assert pol.hash_index_get(h, "text") == prefix
    assert pol.hash_index_get(orc.hash_node_path(texts[0]), "text") is not None
    # evict_cache clears a model's map once it outgrows max_size (:335-351)
    pol.evict_cache(3); op.evict_cache(3)
    assert pol.hash_index_size("tokens") == len(op.hash_index("tokens")) == 0
    assert pol.hash_index_size("text") == len(op.hash_index("text")) == 0
