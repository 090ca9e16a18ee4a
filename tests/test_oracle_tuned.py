"""The "port-tuned" CPU variant used by bench.py's reference arm (oracle/tuned_event.h: flat bitset index, no per-block allocation) must
return exactly the plain oracle's picks — randomized index states incl. mid-sequence removals (non-prefix-closed), Multi entries, unhealthy
workers, duplicate URLs, wide fleets (> 64 workers) and the imbalance gate."""
import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth


@pytest.mark.parametrize("seed,n_workers,T,bs,jump,B", [(1, 64, 512, 16, 64, 400), (2, 64, 512, 16, 8, 300), (3, 200, 512, 16, 4, 300), (4, 10, 96, 4, 4, 300),
                                                        (5, 64, 1024, 32, 16, 200)])
def test_tuned_equals_plain_oracle(seed, n_workers, T, bs, jump, B):
    rng = np.random.default_rng(seed)
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=bs)
    base = synth.worker_urls(n_workers)
    urls = list(base) + [base[1], base[min(5, n_workers - 1)], base[1]]          # duplicate URLs in the slice
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
    op.set_workers(urls)
    oix = orc.PositionalIndexer(jump)
    op.attach_indexer("unknown", oix)
    op.set_block_size("unknown", bs)
    op.set_kv_event_monitor(True)
    for u in (base[i] for i in rng.permutation(n_workers)):
        oix.intern_worker(u)
    seqs = synth.gen_sequences(150, T, seed)
    P, sid = T // bs, 1
    for s in range(len(seqs)):
        hs = orc.compute_request_content_hashes(seqs[s], bs)
        for w in rng.choice(n_workers, size=int(rng.integers(1, 5)), replace=False):
            depth = int(rng.integers(1, P + 1))
            start = int(rng.integers(0, 3)) if rng.random() < 0.2 else 0      # a chain that starts deeper → different prefix hashes → Multi entries
            blocks = [(sid + i, hs[start + i]) for i in range(max(1, depth - start))]
            sid += len(blocks)
            wid = oix.worker_id(base[int(w)])
            oix.apply_stored(wid, blocks)
            if rng.random() < 0.15 and len(blocks) > 2:
                oix.apply_removed(wid, [blocks[int(rng.integers(0, len(blocks)))][0]])
    q = synth.gen_queries(seqs, B, seed, block=bs)
    lens = rng.integers(0, T + 1, size=B); lens[: B // 3] = T
    flat = np.concatenate([q[i, : lens[i]] for i in range(B)]).astype(np.uint32)
    offs = np.zeros(B + 1, np.uint64); np.cumsum(lens, out=offs[1:])
    n = len(urls)
    for rnd in range(4):
        loads = rng.integers(0, 12, size=n)
        if rnd == 3:
            loads[2] += 500                                                     # imbalanced → first-min-load for every request
        healthy = (rng.random(n) > 0.15).astype(np.uint8)
        circuit = (rng.random(n) > 0.05).astype(np.uint8)
        op.set_state(loads, healthy, circuit)
        want, br, _, _ = op.select_batch_tokens(flat, offs)
        for threads in (1, 3):
            got, secs = op.tuned_select_steps_mt(oix, [(flat, offs)], 2, threads, 1.5, 64, bs)
            assert np.array_equal(got, want), (rnd, threads, int((got != want).sum()))
        if rnd < 3:
            assert (np.asarray(br) == 2).sum() > B // 10
