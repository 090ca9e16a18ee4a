"""GPU parity for the HTTP text (string-tree) mode: the reference's string_tree.rs unit tests through the C ABI (host tree
+ GPU walk kernel), the policy-level TEXT_* known answers, BASELINE config 1 (4 workers, 1000 chat-completion routing
texts in the mesh_load_gen shape), and randomized request streams compared decision by decision with the oracle —
one-by-one for the sequential batch mode, as a snapshot batch for the snapshot mode."""
import random

import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth
from tests import scenarios_cache_aware as SC
from tests import scenarios_string_tree as SS

pytestmark = pytest.mark.gpu

CFG = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)


def _mk():
    from smg_b200.policy import Tree
    return Tree.standalone()


@pytest.mark.parametrize("scenario", SS.SCENARIOS, ids=lambda f: f.__name__)
def test_string_tree_scenarios_on_gpu(scenario):
    scenario(_mk)


def _mk_policy(config):
    from smg_b200 import CacheAwarePolicy
    return CacheAwarePolicy(config)


@pytest.mark.parametrize("name", [k for k in SC.ALL if k.startswith("TEXT_")])
def test_text_policy_scenarios_on_gpu(name):
    SC.ALL[name](_mk_policy)


def test_match_epoch_refresh_is_one_in_eight_on_gpu():
    t = _mk()
    t.insert_text("hello", "a")               # epoch 0
    for _ in range(7):
        t.match_prefix_with_counts("hello")   # epochs 1..7: no refresh (string_tree.rs:633-637)
    assert dict(t.entries())["hello"] == [("a", 0)]
    t.match_prefix_with_counts("hello")       # epoch 8 → refresh
    assert dict(t.entries())["hello"] == [("a", 8)]


def test_invalid_utf8_is_rejected():
    import ctypes as C
    from smg_b200 import SmgxError
    t = _mk()
    bad = np.frombuffer(b"ab\xc3", dtype=np.uint8).copy()
    with pytest.raises(SmgxError):
        t.h.call("smgx_stree_insert_text", t.model, bad.ctypes.data_as(C.c_void_p), 3, b"x")


TOPICS = ["Explain", "Summarise", "Translate to 中文:", "Write a haiku about", "Debug this code:", "List facts on", "Écris un poème sur", "🔥 Roast"]
SUBJECTS = ["radix trees", "the B200 memory system", "load balancing", "cache-aware routing", "naïve café crème brûlée", "GPU 调度", "emoji 👋🌍 parsing"]


def _chat_texts(n, seed):
    """Routing texts in the shape of scripts/mesh_load_gen.py:98-108 ("{topic} {subject} {extra} (variant k)"), with a
    shared system prompt in front as extract_text_for_routing (chat.rs:598) concatenates the messages."""
    r = random.Random(seed)
    system = ["You are a helpful assistant.", "You are a terse assistant. Answer in one line.", ""]
    out = []
    for _ in range(n):
        s = r.choice(system)
        body = f"{r.choice(TOPICS)} {r.choice(SUBJECTS)} {'x' * r.randrange(0, 40)} (variant {r.randrange(0, 50)})"
        out.append((s + " " + body) if s else body)
    return out


def _pair(n_workers, mode):
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    urls = synth.worker_urls(n_workers)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), tree_batch_mode=mode)
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    return pol, ws, op, urls


def _check_batch(pol, ws, op, texts, snapshot, tag):
    idx, info = pol.select_worker_batch_request_text(ws, texts)
    if snapshot:
        want, br, ma, inp, _ = op.select_batch_text(texts, snapshot=True)
        for i in range(len(texts)):
            assert want[i] == idx[i] and br[i] == info[i].branch and ma[i] == info[i].matched and inp[i] == info[i].input, (tag, i, texts[i])
    else:
        for i, t in enumerate(texts):
            d = op.select_worker(request_text=t)
            assert (d.idx if d.idx is not None else -1) == idx[i], (tag, i, t, d.branch, orc.BRANCHES[info[i].branch])
            assert d.branch == orc.BRANCHES[info[i].branch]
            assert d.matched == info[i].matched and d.input == info[i].input, (tag, i, t)


@pytest.mark.parametrize("mode", ["sequential", "snapshot"])
def test_config1_http_mode_1000_chat_bodies(mode):
    """BASELINE config 1: 4 workers 127.0.0.1:9000..9003, CLI-default thresholds, loads all 0, 1000 chat routing texts."""
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    urls = [f"http://127.0.0.1:{9000 + i}" for i in range(4)]
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), tree_batch_mode=mode)
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    texts = _chat_texts(1000, 0)
    for b in range(0, 1000, 125):
        _check_batch(pol, ws, op, texts[b:b + 125], mode == "snapshot", b)
    assert pol.string_tree().entries() == op.string_tree().entries()
    assert pol.string_tree().get_tenant_char_count() == op.string_tree().get_tenant_char_count()


@pytest.mark.parametrize("mode", ["sequential", "snapshot"])
@pytest.mark.parametrize("seed,n_workers", [(1, 4), (2, 16), (3, 64)])
def test_random_text_stream_parity(seed, n_workers, mode):
    """Batches with repeated / shared prefixes, changing fleet state, an imbalanced batch and eviction in the middle."""
    pol, ws, op, urls = _pair(n_workers, mode)
    rng = np.random.default_rng(seed)
    for batch_no in range(10):
        loads = rng.integers(0, 12, size=n_workers)
        healthy = (rng.random(n_workers) > 0.08).astype(np.uint8)
        if batch_no == 6:
            loads[0] += 200          # imbalanced batch: min-load picks that still update the tree (cache_aware.rs:403-425)
        for w, l, h in zip(ws, loads, healthy):
            w.set_load(int(l)); w.set_healthy(bool(h))
        op.set_state(loads, healthy, [1] * n_workers)
        texts = _chat_texts(int(rng.integers(1, 60)), 1000 * seed + batch_no)
        if batch_no % 3 == 0:
            texts += ["", "Z", "你", "你好", texts[0], texts[0][: len(texts[0]) // 2]]
        _check_batch(pol, ws, op, texts, mode == "snapshot", batch_no)
        if batch_no == 4:
            pol.evict_cache(200)
            op.evict_cache(200)
            assert pol.string_tree().entries() == op.string_tree().entries()
    assert pol.string_tree().entries() == op.string_tree().entries()          # same nodes, tenants and epochs
    assert pol.string_tree().get_tenant_char_count() == op.string_tree().get_tenant_char_count()
    assert pol.string_tree().get_used_size_per_tenant() == op.string_tree().get_used_size_per_tenant()
    assert pol.string_tree().node_count() == op.string_tree().node_count()
    assert list(pol.take_processed()) == [op.processed(i) for i in range(n_workers)]


def test_snapshot_mode_differs_from_sequential_only_on_intra_batch_conflicts():
    pol_a, ws_a, _, _ = _pair(4, "sequential")
    pol_b, ws_b, _, _ = _pair(4, "snapshot")
    distinct = [f"{chr(ord('a') + i)} unique request number {i}" for i in range(20)]
    ia, fa = pol_a.select_worker_batch_request_text(ws_a, distinct)
    ib, fb = pol_b.select_worker_batch_request_text(ws_b, distinct)
    assert list(ia) == list(ib) and [x.matched for x in fa] == [x.matched for x in fb]
    assert pol_a.string_tree().entries() == pol_b.string_tree().entries()      # timestamps included
    dup = ["Zame text twice in one batch"] * 2   # no earlier text starts with "Z"
    ia, fa = pol_a.select_worker_batch_request_text(ws_a, dup)
    ib, fb = pol_b.select_worker_batch_request_text(ws_b, dup)
    assert fa[1].matched == len(dup[0]) and fb[1].matched == 0                 # the 2nd request sees the 1st insert only sequentially
