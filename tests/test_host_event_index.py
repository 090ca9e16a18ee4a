"""CPU-only: the library's host-side index writer (event_index.cu) against the oracle on random event streams.
Only the writer-visible counters can be compared without a GPU (current_size / entry_count / error codes);
query parity is tests/test_gpu_event_*.py."""
import random

import pytest

from oracle import orc
from smg_b200.policy import ApplyError, CacheAwareConfig, PositionalIndexer, _Handle


@pytest.mark.parametrize("seed,n_workers", [(1, 3), (2, 70), (3, 200)])
def test_random_event_stream_counters(seed, n_workers):
    rng = random.Random(seed)
    h = _Handle(CacheAwareConfig(eviction_interval_secs=0), device_id=-1)
    gpu = PositionalIndexer(h, "unknown", 8)
    ref = orc.PositionalIndexer(8)
    wids = []
    for i in range(n_workers):
        a, b = gpu.intern_worker(f"http://w{i}"), ref.intern_worker(f"http://w{i}")
        assert a == b
        wids.append(a)
    chains = {}   # worker → list of stored (seq, content) in chain order
    next_seq = [1]
    for step in range(1500):
        w = rng.choice(wids)
        op = rng.random()
        if op < 0.55:
            n = rng.randint(1, 6)
            content = [rng.randint(1, 40) for _ in range(n)]           # small alphabet → shared entries, Multi upgrades
            blocks = [(next_seq[0] + i, c) for i, c in enumerate(content)]
            next_seq[0] += n
            parent = None
            if chains.get(w) and rng.random() < 0.6:
                parent = rng.choice(chains[w])[0]
            if rng.random() < 0.05:
                parent = 10**12 + step  # unknown parent → error path
            errs = []
            for ix in (gpu, ref):
                try:
                    ix.apply_stored(w, blocks, parent)
                    errs.append(None)
                except Exception as e:  # noqa: BLE001
                    errs.append("NotTracked" if "WorkerNotTracked" in str(e) else "ParentNotFound" if "ParentBlockNotFound" in str(e) else str(e))
            assert errs[0] == errs[1]
            if errs[0] is None:
                chains.setdefault(w, []).extend(blocks)
        elif op < 0.85:
            if chains.get(w):
                k = rng.randint(1, min(4, len(chains[w])))
                victims = rng.sample(chains[w], k)
                for ix in (gpu, ref):
                    ix.apply_removed(w, [v[0] for v in victims] + [999_999_999])
                chains[w] = [b for b in chains[w] if b not in victims]
        elif op < 0.95:
            for ix in (gpu, ref):
                ix.apply_cleared(w)
            chains[w] = []
        else:
            for ix in (gpu, ref):
                ix.remove_worker(w)
            chains[w] = []
        assert gpu.current_size() == ref.current_size()
        assert gpu.entry_count() == ref.entry_count()
