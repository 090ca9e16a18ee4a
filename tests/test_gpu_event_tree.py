"""GPU parity: the reference's event_tree.rs unit tests (scenarios_event_tree.py) executed through the C ABI —
host writer + find_matches kernel — plus randomized find_matches parity against the oracle, including
non-prefix-closed states, Multi entries and fleets wider than one bitset word."""
import random

import numpy as np
import pytest

from oracle import orc
from tests import scenarios_event_tree as S

pytestmark = pytest.mark.gpu


def _mk(jump):
    from smg_b200.policy import PositionalIndexer
    return PositionalIndexer.standalone(jump)


@pytest.mark.parametrize("scenario", S.SCENARIOS, ids=lambda f: f.__name__)
def test_event_tree_scenarios_on_gpu(scenario):
    scenario(_mk, orc)   # `orc` only supplies the hash helpers used to BUILD the inputs


def test_zero_jump_size_panics():
    with pytest.raises(ValueError, match="jump_size must be greater than 0"):
        _mk(0)


def test_content_hashes_kernel_matches_oracle_all_block_sizes():
    from smg_b200 import _lib
    from smg_b200.policy import CacheAwareConfig, _Handle
    import ctypes as C
    h = _Handle(CacheAwareConfig(eviction_interval_secs=0), 0)
    rng = np.random.default_rng(5)
    for bs in (1, 2, 3, 4, 5, 8, 13, 16, 31, 32, 33, 60, 61, 64, 100, 128, 256, 257):
        n = bs * 7 + (bs // 2)
        toks = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        out = np.zeros(8, np.uint64)
        got = C.c_uint32()
        h.call("smgx_content_hashes", toks.ctypes.data_as(C.c_void_p), n, bs, out.ctypes.data_as(C.c_void_p), 8, C.byref(got))
        want = orc.compute_request_content_hashes(toks, bs)
        assert got.value == len(want) == 7
        assert [int(x) for x in out[:7]] == want, bs
    got = C.c_uint32(99)
    h.call("smgx_content_hashes", None, 3, 0, None, 0, C.byref(got))  # block_size 0 → empty (event_tree.rs:142-145)
    assert got.value == 0


@pytest.mark.parametrize("seed,n_workers,jump", [(11, 5, 4), (12, 64, 8), (13, 65, 3), (14, 300, 16), (15, 40, 1), (16, 130, 64)])
def test_random_find_matches_parity(seed, n_workers, jump):
    """Random store/remove/clear streams over a small content alphabet (shared entries, Single→Multi upgrades,
    mid-sequence gaps), then random queries: per-worker scores and tree sizes must equal the oracle's exactly."""
    rng = random.Random(seed)
    gpu, ref = _mk(jump), orc.PositionalIndexer(jump)
    wids = [gpu.intern_worker(f"http://w{i}") for i in range(n_workers)]
    assert wids == [ref.intern_worker(f"http://w{i}") for i in range(n_workers)]
    base = [[rng.randint(1, 6) for _ in range(40)] for _ in range(6)]   # a few base sequences over a tiny alphabet
    seq_id = [1]
    stored = {w: [] for w in wids}
    for w in wids:
        for _ in range(rng.randint(0, 3)):
            b = rng.choice(base)
            depth = rng.randint(1, 40)
            content = list(b[:depth])
            if rng.random() < 0.3:
                content[rng.randrange(depth)] = rng.randint(1, 6)       # diverge somewhere
            blocks = [(seq_id[0] + i, c) for i, c in enumerate(content)]
            seq_id[0] += depth
            for ix in (gpu, ref):
                ix.apply_stored(w, blocks)
            stored[w].extend(blocks)
        if stored[w] and rng.random() < 0.4:                              # punch holes (no cascade)
            victims = rng.sample(stored[w], rng.randint(1, min(5, len(stored[w]))))
            for ix in (gpu, ref):
                ix.apply_removed(w, [v[0] for v in victims])
        if rng.random() < 0.05:
            for ix in (gpu, ref):
                ix.apply_cleared(w)
    assert gpu.current_size() == ref.current_size() and gpu.entry_count() == ref.entry_count()
    for q in range(60):
        b = rng.choice(base)
        qlen = rng.randint(1, 40)
        query = list(b[:qlen])
        if rng.random() < 0.4:
            query[rng.randrange(qlen)] = rng.randint(1, 6)
        for ee in (False, True):
            assert gpu.find_matches(query, ee) == ref.find_matches(query, ee), (q, ee, query)
