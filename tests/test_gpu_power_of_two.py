"""power_of_two on the GPU (smgx_power_of_two_select_batch) against the oracle: the same draw stream, so picks, candidate pairs and the
metric used are compared bit for bit on random fleets; plus the reference's unit tests through the product's policy object."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu


def _fleet(rng, w):
    urls = [f"http://worker-{i}:8000" for i in range(w)]
    loads = rng.integers(0, 50, size=w)
    healthy = (rng.random(w) > 0.2).astype(np.uint8)
    circuit = (rng.random(w) > 0.1).astype(np.uint8)
    cached = {u: float(rng.random()) for u in urls if rng.random() < 0.6}
    return urls, loads, healthy, circuit, cached


@pytest.mark.parametrize("w,n,seed", [(1, 10, 1), (2, 1000, 2), (3, 5000, 3), (17, 20000, 4), (64, 100000, 5), (512, 65536, 6), (4096, 30000, 7)])
def test_random_fleets_equal_oracle(w, n, seed):
    from smg_b200 import BasicWorker, PowerOfTwoPolicy
    rng = np.random.default_rng(seed)
    pol = PowerOfTwoPolicy(max_batch=max(n, 4096))
    opol = orc.PowerOfTwoPolicy()
    for rnd in range(3):
        urls, loads, healthy, circuit, cached = _fleet(rng, w)
        ws = [BasicWorker(u) for u in urls]
        for i, x in enumerate(ws):
            x.set_load(int(loads[i])); x.set_healthy(bool(healthy[i])); x.set_circuit_ok(bool(circuit[i]))
        pol.update_loads(cached); opol.update_loads(cached)
        s = int(rng.integers(0, 2**63))
        idx, pairs, metric = pol.select_worker_batch(ws, n, seed=s, with_details=True)
        oidx, opairs, ometric, oproc = opol.select_batch(urls, loads, healthy, circuit, n, s)
        assert np.array_equal(idx, oidx) and np.array_equal(pairs, opairs) and np.array_equal(metric, ometric)
        if (healthy & circuit).sum() > 1:
            assert np.array_equal(np.asarray([x.processed() for x in ws], np.uint64), oproc)


def test_reference_unit_tests_through_the_product():
    """power_of_two.rs:169-420 on the product's policy object (its own seed sequence)."""
    from smg_b200 import BasicWorker, PolicyFactory

    def mk(url, reqs=0):
        w = BasicWorker(url)
        w.set_load(reqs)
        return w
    pol = PolicyFactory.create_by_name("power_of_two")
    assert pol.name() == "power_of_two"
    ws = [mk("http://w1:8000", 10), mk("http://w2:8000", 5), mk("http://w3:8000", 0)]
    counts = np.bincount([pol.select_worker(ws) for _ in range(100)], minlength=3)                    # test_power_of_two_selection
    assert counts[2] > counts[1] > counts[0]
    pol = PolicyFactory.create_by_name("PowerOfTwo")
    ws = [mk("http://w1:8000"), mk("http://w2:8000")]
    pol.update_loads({"http://w1:8000": [0.8], "http://w2:8000": [0.1]})                            # make_load(x): one DP rank with token_usage x
    assert sum(pol.select_worker(ws) == 1 for _ in range(50)) > 35                                    # test_power_of_two_with_cached_loads
    assert PolicyFactory.create_by_name("power_of_two").select_worker([mk("http://w1:8000")]) == 0   # test_power_of_two_single_worker
    pol = PolicyFactory.create_by_name("power_of_two")
    pol.update_loads({"http://worker_a:8000": [0.9]})
    assert pol.select_worker([mk("http://worker_a:8000"), mk("http://worker_b:8000", 5)]) == 0       # test_reproduce_incompatible_metric_bug
    pol = PolicyFactory.create_by_name("power_of_two")                                                # test_power_of_two_edge_cases
    pol.update_loads({"http://a:8000": [0.1], "http://b:8000": [0.9]})
    assert pol.select_worker([mk("http://a:8000", 10), mk("http://b:8000", 2)]) == 0
    pol.update_loads({"http://c:8000": [0.1]})
    assert pol.select_worker([mk("http://c:8000", 10), mk("http://d:8000", 2)]) == 1
    pol.update_loads({"http://f:8000": [0.1]})
    assert pol.select_worker([mk("http://e:8000", 2), mk("http://f:8000", 10)]) == 0
    pol.update_loads({})
    assert pol.select_worker([mk("http://g:8000", 5), mk("http://h:8000", 3)]) == 1
    unhealthy = [mk("http://x:8000"), mk("http://y:8000")]
    for w in unhealthy:
        w.set_healthy(False)
    assert pol.select_worker(unhealthy) is None
