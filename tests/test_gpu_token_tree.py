"""GPU parity for the approximate token-tree mode: the reference's token_tree.rs unit tests through the C ABI (host tree +
GPU match kernel), the policy-level TOKEN_* known answers, and randomized request streams compared decision by decision
with the oracle run sequentially (the reference's semantics) — batched on the GPU side."""
import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth
from tests import scenarios_cache_aware as SC
from tests import scenarios_token_tree as ST

pytestmark = pytest.mark.gpu


def _mk(policy):
    from smg_b200 import TokenTree
    return TokenTree.standalone(policy)


@pytest.mark.parametrize("scenario", ST.SCENARIOS, ids=lambda f: f.__name__)
def test_token_tree_scenarios_on_gpu(scenario):
    scenario(_mk)


def _mk_policy(config):
    from smg_b200 import CacheAwarePolicy
    return CacheAwarePolicy(config)


@pytest.mark.parametrize("name", [k for k in SC.ALL if k.startswith(("TOKEN_", "EVENT_"))])
def test_policy_scenarios_on_gpu(name):
    SC.ALL[name](_mk_policy)


@pytest.mark.parametrize("mode", ["sequential", "snapshot"])
@pytest.mark.parametrize("seed,n_workers,shape", [(1, 8, "trunks"), (2, 64, "trunks"), (3, 16, "zipf"), (4, 5, "ragged")])
def test_random_stream_parity_with_oracle(seed, n_workers, shape, mode):
    """Batches of requests (with repeated / shared prefixes inside a batch) routed through the GPU path must equal the
    oracle with the same frozen fleet snapshot — called one request at a time (sequential batch mode, the reference's
    one-by-one semantics) or as one snapshot batch (snapshot mode): pick, branch and matched tokens."""
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    rng = np.random.default_rng(seed)
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)
    urls = synth.worker_urls(n_workers)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg), tree_batch_mode=mode)
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
    op.set_workers(urls)

    trunks = [rng.integers(0, 50000, size=16 * int(rng.integers(1, 6)), dtype=np.uint32) for _ in range(12)]
    zw = 1.0 / np.arange(1, 13) ** 1.1
    zw /= zw.sum()

    def make_request():
        if shape == "zipf":
            t = trunks[int(rng.choice(12, p=zw))]
        else:
            t = trunks[int(rng.integers(0, 12))]
        k = int(rng.integers(0, 4))
        tail = rng.integers(0, 50000, size=16 * k + (int(rng.integers(0, 16)) if shape == "ragged" else 0), dtype=np.uint32)
        cut = len(t) if rng.random() < 0.7 else 16 * int(rng.integers(0, len(t) // 16 + 1))
        req = np.concatenate([t[:cut], tail])
        if shape == "ragged" and rng.random() < 0.1:
            req = req[: int(rng.integers(0, 15))]
        return req

    for batch_no in range(12):
        loads = rng.integers(0, 12, size=n_workers)
        healthy = (rng.random(n_workers) > 0.08).astype(np.uint8)
        if batch_no == 7:
            loads[0] += 200          # imbalanced batch: min-load picks that still update the tree
        for w, l, h in zip(ws, loads, healthy):
            w.set_load(int(l)); w.set_healthy(bool(h))
        op.set_state(loads, healthy, [1] * n_workers)
        reqs = [make_request() for _ in range(int(rng.integers(1, 70)))]
        idx, info = pol.select_worker_batch(ws, reqs)
        if mode == "snapshot":
            flat = np.concatenate(reqs).astype(np.uint32) if sum(len(r) for r in reqs) else np.zeros(0, np.uint32)
            off = np.zeros(len(reqs) + 1, np.uint64)
            np.cumsum([len(r) for r in reqs], out=off[1:])
            want, br, ma, _ = op.select_batch_tokens(flat, off, snapshot=True)
            for i in range(len(reqs)):
                assert want[i] == idx[i] and br[i] == info[i].branch and ma[i] == info[i].matched, (batch_no, i)
            continue
        for i, r in enumerate(reqs):
            d = op.select_worker(tokens=r)
            assert (d.idx if d.idx is not None else -1) == idx[i], (batch_no, i, d.branch, orc.BRANCHES[info[i].branch])
            assert d.branch == orc.BRANCHES[info[i].branch]
            assert d.matched == info[i].matched and d.input == info[i].input
    # the two trees ended up identical: same entries, same timestamps
    got = pol.token_tree().entries()
    want = op.token_tree().entries()
    assert got == want
    for u in urls:
        assert pol.token_tree().tenant_token_size(u) == op.token_tree().tenant_token_size(u)
    # eviction keeps them identical
    pol.evict_cache(64)
    op.evict_cache(64)
    assert pol.token_tree().entries() == op.token_tree().entries()
