"""The CPU oracle of PowerOfTwoPolicy against the reference's own unit tests (model_gateway/src/policies/power_of_two.rs:169-420, names
kept).  The reference draws from an unseedable thread-local generator, so its tests are statistical or arranged so that the draw cannot
matter; they hold here for every seed tried.  Plus the properties of the draw procedure itself (:49-52)."""
import numpy as np
import pytest

from oracle import orc

W3 = ["http://w1:8000", "http://w2:8000", "http://w3:8000"]


@pytest.mark.parametrize("seed", [1, 2, 3, 12345, 2**63 + 7])
def test_power_of_two_selection(seed):                                 # :169-206
    pol = orc.PowerOfTwoPolicy()
    idx, pairs, metric, proc = pol.select_batch(W3, [10, 5, 0], [1, 1, 1], None, 100, seed)
    counts = np.bincount(idx, minlength=3)
    assert counts[2] > counts[1] > counts[0]
    assert (metric == 0).all()                                          # no cached loads: request counts
    assert np.array_equal(proc, counts)                                 # increment_processed on every pick (:110)


@pytest.mark.parametrize("seed", [1, 2, 3, 99])
def test_power_of_two_with_cached_loads(seed):                         # :210-246
    pol = orc.PowerOfTwoPolicy()
    pol.update_loads({"http://w1:8000": 0.8, "http://w2:8000": 0.1})
    idx, _, metric, _ = pol.select_batch(W3[:2], [0, 0], [1, 1], None, 50, seed)
    assert (idx == 1).sum() > 35 and (metric == 1).all()              # with two workers the pair is always {w1, w2}: w2 wins every time
    assert (idx == 1).all()


def test_power_of_two_single_worker():                                 # :249-260
    pol = orc.PowerOfTwoPolicy()
    idx, pairs, metric, proc = pol.select_batch(W3[:1], [0], [1], None, 4, 7)
    assert (idx == 0).all() and (pairs == -1).all() and (metric == 2).all()
    assert proc[0] == 0                                                 # the early return (:44-46) does not reach increment_processed


@pytest.mark.parametrize("seed", range(8))
def test_reproduce_incompatible_metric_bug(seed):                      # :267-320: only A has token data → request counts for BOTH
    pol = orc.PowerOfTwoPolicy()
    pol.update_loads({"http://worker_a:8000": 0.9})
    idx, _, metric, _ = pol.select_batch(["http://worker_a:8000", "http://worker_b:8000"], [0, 5], [1, 1], None, 1, seed)
    assert idx[0] == 0 and metric[0] == 0


@pytest.mark.parametrize("seed", range(8))
def test_power_of_two_edge_cases(seed):                                # :327-420 (one policy object through all four steps)
    pol = orc.PowerOfTwoPolicy()
    pol.update_loads({"http://a:8000": 0.1, "http://b:8000": 0.9})
    assert pol.select_batch(["http://a:8000", "http://b:8000"], [10, 2], [1, 1], None, 1, seed)[0][0] == 0     # both have data: token usage
    pol.update_loads({"http://c:8000": 0.1})
    assert pol.select_batch(["http://c:8000", "http://d:8000"], [10, 2], [1, 1], None, 1, seed)[0][0] == 1     # only c: request counts
    pol.update_loads({"http://f:8000": 0.1})
    assert pol.select_batch(["http://e:8000", "http://f:8000"], [2, 10], [1, 1], None, 1, seed)[0][0] == 0     # only f: request counts
    pol.update_loads({})
    assert pol.select_batch(["http://g:8000", "http://h:8000"], [5, 3], [1, 1], None, 1, seed)[0][0] == 1      # none: request counts


def test_effective_token_usage():                                       # protocols worker.rs:1039-1044
    assert orc.PowerOfTwoPolicy.effective_token_usage([]) == 0.0
    assert orc.PowerOfTwoPolicy.effective_token_usage([0.2, 0.4, 0.9]) == (0.2 + 0.4 + 0.9) / 3


def test_no_healthy_and_circuit_breaker():                             # :40-42, policies/mod.rs:137-144
    pol = orc.PowerOfTwoPolicy()
    assert (pol.select_batch(W3, [1, 2, 3], [0, 0, 0], None, 5, 1)[0] == -1).all()
    idx, pairs, _, _ = pol.select_batch(W3, [9, 0, 0], [1, 1, 1], [1, 0, 0], 5, 1)                             # only w1 passes both checks
    assert (idx == 0).all() and (pairs == -1).all()
    idx, pairs, _, _ = pol.select_batch(W3, [9, 1, 0], [1, 0, 1], [1, 1, 1], 64, 3)
    assert set(idx.tolist()) == {2} and set(pairs.flatten().tolist()) == {0, 2}                               # w2 is never a candidate


@pytest.mark.parametrize("h", [2, 3, 7, 64])
def test_draw_procedure(h):
    """idx2 = (idx1 + 1 + uniform(0..h-1)) % h is never idx1, every ordered pair is reachable, candidates are uniform; ties keep the first."""
    pol = orc.PowerOfTwoPolicy()
    urls = [f"http://w{i}:8000" for i in range(h)]
    n = 20000
    idx, pairs, _, _ = pol.select_batch(urls, [0] * h, [1] * h, None, n, 0xC0FFEE)
    assert (pairs[:, 0] != pairs[:, 1]).all()
    assert np.array_equal(idx, pairs[:, 0])                             # equal loads: load1 <= load2 keeps candidate 1 (:91-95)
    c1, c2 = np.bincount(pairs[:, 0], minlength=h), np.bincount(pairs[:, 1], minlength=h)
    assert c1.min() > 0.8 * n / h and c1.max() < 1.2 * n / h and c2.min() > 0.8 * n / h and c2.max() < 1.2 * n / h
    if h <= 7:
        assert len({(a, b) for a, b in pairs.tolist()}) == h * (h - 1)
    # a different seed is a different stream; the same seed the same picks
    again = pol.select_batch(urls, [0] * h, [1] * h, None, n, 0xC0FFEE)[1]
    other = pol.select_batch(urls, [0] * h, [1] * h, None, n, 0xC0FFEF)[1]
    assert np.array_equal(again, pairs) and not np.array_equal(other, pairs)
