"""Load-feedback batches (smgx_set_load_feedback): "request i sees the fleet snapshot plus the picks of requests < i" — the reference's
request stream, where the router takes a WorkerLoadGuard on the chosen worker right after every select_worker call
(routers/http/router.rs:319-321, worker/worker.rs:1067-1070).  Oracle = select_worker called one request at a time with the chosen
worker's load bumped after each pick (orc_policy_select_batch_tokens_feedback)."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(n_workers, T, bs, jump, seed, n_seq=300):
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=bs)
    urls = synth.worker_urls(n_workers)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg), max_batch=8192, max_tokens_per_request=T)
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
    op.set_workers(urls)
    oix = orc.PositionalIndexer(jump)
    op.attach_indexer("unknown", oix)
    op.set_block_size("unknown", bs)
    op.set_kv_event_monitor(True)
    rng = np.random.default_rng(seed)
    for u in urls:
        assert ix.intern_worker(u) == oix.intern_worker(u)
    seqs = synth.gen_sequences(n_seq, T, seed)
    P, sid = T // bs, 1
    for s in range(n_seq):
        hs = orc.compute_request_content_hashes(seqs[s], bs)
        for w in rng.choice(n_workers, size=int(rng.integers(1, 4)), replace=False):   # 1-3 workers per sequence → tied overlap sets
            depth = P if rng.random() < 0.7 else int(rng.integers(1, P + 1))
            blocks = [(sid + i, hs[i]) for i in range(depth)]
            sid += depth
            ix.apply_stored(int(w), blocks); oix.apply_stored(int(w), blocks)
    return pol, ws, op, seqs, rng


@pytest.mark.parametrize("n_workers,seed", [(64, 1), (64, 2), (256, 3)])
def test_feedback_stream_matches_one_by_one_reference(n_workers, seed):
    T, bs, B = 512, 16, 4096
    pol, ws, op, seqs, rng = _setup(n_workers, T, bs, 64, seed)
    pol.set_load_feedback(True)
    # 70 % cached sequences, 30 % novel (→ the min-load branch): far more than 10 % of the batch takes the load-dependent path
    q = synth.gen_queries(seqs, B, seed, block=bs)
    novel = rng.random(B) < 0.3
    q[novel] = rng.integers(0, 50000, size=(int(novel.sum()), T), dtype=np.uint32)
    tokens, offsets = synth.ragged(q)
    for rnd in range(3):
        loads = synth.poisson_loads(n_workers, 8, seed + rnd)
        healthy = (rng.random(n_workers) > 0.1).astype(np.uint8)
        if rnd == 2:
            loads[:] = 3
            healthy[:] = 1          # (min / max run over ALL workers: an unhealthy one would pin the minimum and keep the gate open)
            loads[5] = 80           # starts imbalanced: the gate sends everything to the least loaded worker until max - min <= 64 (≈ 800 picks)
        for i, w in enumerate(ws):
            w.set_load(int(loads[i])); w.set_healthy(bool(healthy[i]))
        op.set_state(loads, healthy, [1] * n_workers)
        idx, info = pol.select_worker_batch(ws, tokens=tokens, offsets=offsets)
        oidx, obr, oma, oloads = op.select_batch_tokens_feedback(tokens, offsets.astype(np.uint64))
        assert np.array_equal(idx, oidx), f"round {rnd}: {(idx != oidx).sum()} picks differ"
        assert [i.branch for i in info] == list(obr)
        # overlap scores are reported in blocks, the imbalanced branch's tree match (select_worker_min_load) in tokens
        assert [i.matched * bs if i.branch == 2 else i.matched for i in info] == list(oma)
        minload = np.isin(np.asarray(obr), (1, 3))                   # load-dependent picks: imbalanced-gate and no-overlap min-load
        assert minload.sum() > B // 10
        assert len(set(int(x) for x in oidx[minload])) > 8          # water-filling: the min-load picks spread over the fleet …
        if rnd == 2:
            assert (np.asarray(obr) == 1).sum() > 100 and (np.asarray(obr) == 2).sum() > 0   # the gate was on, then closed inside the batch
    # … whereas the frozen snapshot sends every min-load request of a batch to ONE worker
    pol.set_load_feedback(False)
    idx2, info2 = pol.select_worker_batch(ws, tokens=tokens, offsets=offsets)
    fr = np.asarray([i.branch for i in info2])
    if (fr == 3).sum():
        assert len(set(int(x) for x in idx2[fr == 3])) == 1


def test_feedback_across_batches_of_one_call():
    """Several device-resident batches handed over in one smgx_select_many_tokens_device call form ONE stream."""
    from smg_b200 import _lib
    T, bs, B, NB = 256, 16, 1000, 3
    pol, ws, op, seqs, rng = _setup(64, T, bs, 64, 7, n_seq=150)
    pol.set_load_feedback(True)
    loads = synth.poisson_loads(64, 8, 7)
    for i, w in enumerate(ws):
        w.set_load(int(loads[i]))
    op.set_state(loads, [1] * 64, [1] * 64)
    h, L = pol._h, _lib.load()
    model = pol._push_fleet(ws)
    err = _lib.new_err()
    offsets = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
    d_off = L.smgx_device_alloc(h.p, offsets.nbytes, C.byref(err))
    h.call("smgx_memcpy_h2d", d_off, offsets.ctypes.data_as(C.c_void_p), offsets.nbytes)
    host, d_tok, d_out = [], [], []
    for r in range(NB):
        q = synth.gen_queries(seqs, B, 70 + r, block=bs)
        nov = rng.random(B) < 0.3
        q[nov] = rng.integers(0, 50000, size=(int(nov.sum()), T), dtype=np.uint32)
        flat = np.ascontiguousarray(q.reshape(-1))
        host.append(flat)
        dt = L.smgx_device_alloc(h.p, flat.nbytes, C.byref(err))
        h.call("smgx_memcpy_h2d", dt, flat.ctypes.data_as(C.c_void_p), flat.nbytes)
        d_tok.append(dt); d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
    h.call("smgx_select_many_tokens_device", model, NB, (C.c_void_p * NB)(*d_tok), (C.c_void_p * NB)(*[d_off] * NB), (C.c_uint32 * NB)(*[B] * NB), T, (C.c_void_p * NB)(*d_out))
    h.call("smgx_synchronize")
    want = op.select_batch_tokens_feedback(np.concatenate(host), (np.arange(NB * B + 1, dtype=np.uint64) * T))[0]
    got = np.zeros(B, np.int32)
    for r in range(NB):
        h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[r], B * 4)
        assert np.array_equal(got, want[r * B:(r + 1) * B]), r
    for d in d_tok + d_out + [d_off]:
        L.smgx_device_free(h.p, d)
