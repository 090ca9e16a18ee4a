"""GPU parity at the scale the headline number is measured at:
  * BASELINE config 2 exactly — 64 workers, 31 250 sequences × 32 blocks = 1.0 M index entries (4 M-slot table), batches of 4096
    512-token requests, 37 batches handed to smgx_select_many_tokens_device in one call (→ a 32-batch launch and a 5-batch launch on two
    stream lanes) — every pick compared with the oracle;
  * both implementations of the event-driven pick (fused one-kernel path, round-1 hash + search pair) on the randomized small cases;
  * duplicate URLs in the worker slice in event-driven mode (two slice entries share one indexer id; score_overlap scores both)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth

pytestmark = pytest.mark.gpu

PATHS = {"split": 0, "hs": 2, "stream": 3}     # smgx_set_event_path; everything else is a variant of path 1

CFG = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)


@pytest.fixture
def event_path():
    from smg_b200 import _lib
    L = _lib.load()
    def setter(fused, minb=0, pf=-1, tile=None, min_total=-1, simple=0, depth=0):
        L.smgx_set_event_path(int(fused), minb)      # 0 = pair, 1 (True) = warp-per-request family, 2 = hash stream + last-arriver search, 3 = persistent streaming kernel
        L.smgx_set_event_simple(simple)
        L.smgx_set_tile_depth(depth)
        if pf >= 0:
            L.smgx_set_fused_prefetch(pf)
        if tile is not None:
            L.smgx_set_fused_tile(tile, min_total)
    yield setter
    L.smgx_set_event_path(0, 4)   # the default: the pair (hash stream + balanced search)
    L.smgx_set_fused_prefetch(1)
    L.smgx_set_fused_tile(16, -1)
    L.smgx_set_event_simple(5)
    L.smgx_set_tile_depth(0)


def _config2(n_seq, W, T, bs, B):
    import bench
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), device_id=0, max_batch=B, max_tokens_per_request=T)
    urls = synth.worker_urls(W)
    ws = [BasicWorker(u) for u in urls]
    loads = synth.poisson_loads(W, 8, 42)
    for w, l in zip(ws, loads):
        w.set_load(int(l))
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", 64)
    pol.set_kv_event_monitor(mon)
    seqs = bench.build_population(n_seq, T, 42)
    hashes = bench.populate(pol, ix, seqs, W, bs)
    o_hashes = bench.oracle_hashes(seqs, bs)
    assert np.array_equal(hashes, o_hashes)                 # GPU content hashes of the whole population = the oracle's
    op, _ = bench.populate_oracle(seqs, o_hashes, W, bs, 64)
    op.set_state(loads, [1] * W, [1] * W)
    return pol, ws, ix, op, seqs


@pytest.mark.parametrize("variant", ["stream", "hs", "split", "simple", "tile16", "tile8", "tile32", "tile16-m3", "tile16-d4", "tile8-d8", "fused4", "fused3", "fused4-pf2", "fused4-pf0"])
def test_config2_full_scale_multi_launch(variant, event_path):
    import bench
    from smg_b200 import _lib
    tile = int(variant[4:6].rstrip("-")) if variant.startswith("tile") else 0
    event_path(PATHS.get(variant, 1), 3 if variant.endswith("3") else 4, 2 if variant.endswith("pf2") else 0 if variant.endswith("pf0") else 1, tile=tile,
               simple=5 if variant == "simple" else 0, depth=4 if "-d4" in variant else 8 if "-d8" in variant else 0)
    n_seq, W, T, bs, B, NB = 31250, 64, 512, 16, 4096, 37
    pol, ws, ix, op, seqs = _config2(n_seq, W, T, bs, B)
    assert ix.entry_count() == n_seq * (T // bs)
    h, L = pol._h, _lib.load()
    model = pol._push_fleet(ws)
    err = _lib.new_err()
    offsets = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
    d_off = L.smgx_device_alloc(h.p, offsets.nbytes, C.byref(err))
    h.call("smgx_memcpy_h2d", d_off, offsets.ctypes.data_as(C.c_void_p), offsets.nbytes)
    host, d_tok, d_out = [], [], []
    for r in range(NB):
        flat = np.ascontiguousarray(bench.gen_batch(seqs, B, 900 + r, bs)[0].reshape(-1))
        host.append(flat)
        dt = L.smgx_device_alloc(h.p, flat.nbytes, C.byref(err))
        h.call("smgx_memcpy_h2d", dt, flat.ctypes.data_as(C.c_void_p), flat.nbytes)
        d_tok.append(dt)
        d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
    TOK = (C.c_void_p * NB)(*d_tok); OFF = (C.c_void_p * NB)(*[d_off] * NB); OUT = (C.c_void_p * NB)(*d_out); NS = (C.c_uint32 * NB)(*[B] * NB)
    h.call("smgx_select_many_tokens_device", model, NB, TOK, OFF, NS, T, OUT)
    h.call("smgx_synchronize")
    got = np.zeros(B, np.int32)
    off64 = offsets.astype(np.uint64)
    n_overlap = 0
    for r in range(NB):
        h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[r], B * 4)
        want, br, _, _ = op.select_batch_tokens(host[r], off64)
        assert np.array_equal(got, want), f"batch {r}: {(got != want).sum()} of {B} picks differ from the oracle"
        n_overlap += int((np.asarray(br) == 2).sum())
    assert n_overlap > 0.8 * NB * B        # ≈ 90 % of the mix has a stored prefix
    if variant in PATHS:   # 150 batches in one call: five 32-batch chunks, the hash-scratch ring of the two-lane pipeline wraps around
        rep = [i % NB for i in range(150)]
        h.call("smgx_select_many_tokens_device", model, 150, (C.c_void_p * 150)(*[d_tok[i] for i in rep]), (C.c_void_p * 150)(*[d_off] * 150),
               (C.c_uint32 * 150)(*[B] * 150), T, (C.c_void_p * 150)(*[d_out[i] for i in rep]))
        h.call("smgx_synchronize")
        for r in (0, 17, 36):
            h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[r], B * 4)
            assert np.array_equal(got, op.select_batch_tokens(host[r], off64)[0])
    # a ragged batch with non-uniform sizes per launch (different n per batch → the per-batch lookup path of the fused kernel)
    ns = [4096, 1000, 1, 2500]
    NSr = (C.c_uint32 * 4)(*ns)
    h.call("smgx_select_many_tokens_device", model, 4, (C.c_void_p * 4)(*d_tok[:4]), (C.c_void_p * 4)(*[d_off] * 4), NSr, T, (C.c_void_p * 4)(*d_out[:4]))
    h.call("smgx_synchronize")
    for r, n in enumerate(ns):
        h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[r], n * 4)
        want = op.select_batch_tokens(host[r][: n * T], off64[: n + 1])[0]
        assert np.array_equal(got[:n], want)
    for d in d_tok + d_out + [d_off]:
        L.smgx_device_free(h.p, d)


@pytest.mark.parametrize("case", [(1, 64, 512, 16, 64, 512), (3, 256, 1024, 16, 64, 256), (5, 64, 512, 64, 64, 256), (7, 100, 2048, 32, 32, 128),
                                  (8, 64, 8192, 16, 64, 48)])
@pytest.mark.parametrize("variant", ["split", "hs", "simple", "fused4", "fused3", "tile8", "tile16", "tile32", "tile16-d4"])
def test_random_parity_other_variants(case, variant, event_path):
    """The randomized ragged parity cases of test_gpu_event_select.py (which run the default path: hash stream + balanced search kernel) on every
    other implementation of the event-driven pick."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gpu_event_select", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_event_select.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    test_random_select_parity = mod.test_random_select_parity
    if variant in PATHS:
        event_path(PATHS[variant])
    elif variant.startswith("tile"):   # force the tiled kernel wherever the launch is eligible (≤ 32 blocks, one jump, ≤ 64 workers), however small
        event_path(True, 4, tile=int(variant[4:6].rstrip("-")), min_total=1, depth=4 if "-d4" in variant else 0)
    else:
        event_path(True, 3 if variant == "fused3" else 4, tile=0, simple=5 if variant == "simple" else 0)
    test_random_select_parity(*case)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_duplicate_urls_event_mode(seed, event_path):
    """Two (or three) slice entries with the SAME url map to one indexer id: score_overlap (cache_aware.rs:795-818) scores every healthy
    index, so among duplicates the lowest load wins and, on equal loads, the LAST index; unhealthy duplicates drop out."""
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    rng = np.random.default_rng(seed)
    bs, T, jump, B = 16, 256, 8, 400
    base = synth.worker_urls(24)
    urls = list(base) + [base[3], base[7], base[3], base[20], base[7]]       # duplicates at the end …
    urls.insert(5, base[20])                                                 # … and in the middle
    n = len(urls)
    cfg = dict(CFG)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg))
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
    op.set_workers(urls)
    oix = orc.PositionalIndexer(jump)
    op.attach_indexer("unknown", oix)
    op.set_kv_event_monitor(True)
    for u in base:
        assert ix.intern_worker(u) == oix.intern_worker(u)
    seqs = synth.gen_sequences(120, T, seed)
    P = T // bs
    sid = 1
    for s in range(len(seqs)):
        hs = orc.compute_request_content_hashes(seqs[s], bs)
        for w in rng.choice(len(base), size=int(rng.integers(1, 4)), replace=False):
            depth = int(rng.integers(1, P + 1))
            blocks = [(sid + i, hs[i]) for i in range(depth)]
            sid += depth
            ix.apply_stored(int(w), blocks); oix.apply_stored(int(w), blocks)
    q = synth.gen_queries(seqs, B, seed, block=bs)
    tokens, offsets = synth.ragged(q)
    for variant in ("stream", "hs", "split", "simple", "fused"):
        event_path(PATHS.get(variant, 1), simple=5 if variant == "simple" else 0, tile=0)
        for rnd in range(4):
            loads = rng.integers(0, 6, size=n)            # small range → many equal loads among duplicates
            healthy = (rng.random(n) > 0.2).astype(np.uint8)
            circuit = (rng.random(n) > 0.1).astype(np.uint8)
            for i, w in enumerate(ws):
                w.set_load(int(loads[i])); w.set_healthy(bool(healthy[i])); w.set_circuit_ok(bool(circuit[i]))
            op.set_state(loads, healthy, circuit)
            idx, info = pol.select_worker_batch(ws, tokens=tokens, offsets=offsets)
            oidx, obr, oma, _ = op.select_batch_tokens(tokens, offsets.astype(np.uint64))
            assert np.array_equal(idx, oidx), (variant, rnd)
            assert [i.branch for i in info] == list(obr)
            dup_slices = {i for i, u in enumerate(urls) if urls.count(u) > 1}
            assert len(dup_slices & set(int(x) for x in oidx)) > 0      # duplicates really get picked
