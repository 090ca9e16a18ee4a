#!/usr/bin/env python
"""Secondary measurements (not the headline bench line): approximate token-tree mode and the worker-id-sharded event mode.
Writes one JSON object per mode to stdout.  Run on the GPU box:

    python tools/bench_modes.py tree                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_modes.py sharded
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CFG = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)


def tree_population(n_trunks, rng):
    """SURVEY §8d config-2 tree shape, scaled: trunks of 8 pages × 32 branches of 8 pages × 32 leaves of 16 pages (512-token paths)."""
    paths = []
    for _ in range(n_trunks):
        trunk = rng.integers(0, 50000, size=128, dtype=np.uint32)
        for _b in range(32):
            branch = rng.integers(0, 50000, size=128, dtype=np.uint32)
            for _l in range(32):
                paths.append(np.concatenate([trunk, branch, rng.integers(0, 50000, size=256, dtype=np.uint32)]))
    return paths


def tree_mode():
    from oracle import orc
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy, synth
    W, B, n_trunks = 64, 4096, int(os.environ.get("TRUNKS", "20"))
    rng = np.random.default_rng(42)
    paths = tree_population(n_trunks, rng)
    urls = synth.worker_urls(W)
    bmode = os.environ.get("TREE_BATCH_MODE", "sequential")
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), max_batch=B, max_tokens_per_request=512, tree_batch_mode=bmode)
    ws = [BasicWorker(u) for u in urls]
    for w, l in zip(ws, synth.poisson_loads(W, 8, 42)):
        w.set_load(int(l))
    pol.init_workers(ws)
    tree = pol.token_tree()
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    op.set_state([w.load() for w in ws], [1] * W, [1] * W)
    ot = op.token_tree()
    t0 = time.time()
    for i, p in enumerate(paths):
        tree.insert_tokens(p, urls[i % W])
    t_build = time.time() - t0
    for i, p in enumerate(paths):
        ot.insert_tokens(p, urls[i % W])

    def batch(seed):
        r = np.random.default_rng(seed)
        reqs = []
        for _ in range(B):
            u = r.random()
            p = paths[int(r.integers(0, len(paths)))]
            if u < 0.8:
                reqs.append(p)
            elif u < 0.9:
                k = 16 * int(r.integers(1, 32))
                reqs.append(np.concatenate([p[:k], r.integers(0, 50000, size=512 - k, dtype=np.uint32)]))
            else:
                reqs.append(r.integers(0, 50000, size=512, dtype=np.uint32))
        flat = np.concatenate(reqs).astype(np.uint32)
        return flat, (np.arange(B + 1, dtype=np.uint64) * 512).astype(np.uint32)

    batches = [batch(100 + i) for i in range(4)]
    # parity of the first batch, then timing
    idx, _ = pol.select_worker_batch(ws, tokens=batches[0][0], offsets=batches[0][1])
    snap = bmode == "snapshot"
    want, _, _, t_cpu0 = op.select_batch_tokens(batches[0][0], batches[0][1].astype(np.uint64), snapshot=snap)
    assert np.array_equal(idx, want), "tree-mode picks differ from the oracle"
    t0 = time.perf_counter()
    for b in batches[1:]:
        pol.select_worker_batch(ws, tokens=b[0], offsets=b[1], want_info=False)
    t_gpu = time.perf_counter() - t0
    t_cpu = sum(op.select_batch_tokens(b[0], b[1].astype(np.uint64), snapshot=snap)[3] for b in batches[1:])
    n = 3 * B
    assert pol.token_tree().entries() == op.token_tree().entries(), "trees diverged"
    return {"mode": "approximate token tree (cache_aware.rs:834-904): GPU walk+pick, host-side touches+inserts in request order", "tree_batch_mode": bmode,
            "workers": W, "batch": B, "tree_paths": len(paths), "tree_nodes_approx": n_trunks * (1 + 32 + 1024), "tree_build_s": round(t_build, 2),
            "smgx_decisions_per_s": n / t_gpu, "oracle_1core_decisions_per_s": n / t_cpu, "parity_first_batch": True,
            "kernel_launches": pol.kernel_launches()}


def treewalk_mode():
    """K2a by itself at BASELINE config-2 scale: read-only walk + pick of HBM-resident batches against a ≈1.06 M-node tree
    (TRUNKS trunks of 8 pages × 32 branches of 8 pages × 32 leaves of 16 pages; labels ≈1.07 GB at TRUNKS=1000)."""
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy, _lib, synth
    W, B, T = 64, int(os.environ.get("BATCH", "4096")), 512
    n_trunks = int(os.environ.get("TRUNKS", "1000"))
    ring, steps, per_call = int(os.environ.get("RING", "32")), int(os.environ.get("STEPS", "20")), int(os.environ.get("PER_CALL", "8"))
    rng = np.random.default_rng(42)
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), max_batch=B, max_tokens_per_request=T)
    ws = [BasicWorker(u) for u in urls]
    for w, l in zip(ws, synth.poisson_loads(W, 8, 42)):
        w.set_load(int(l))
    pol.init_workers(ws)
    model = pol._push_fleet(ws)
    h = pol._h
    t0 = time.time()
    sample = []
    url_c = [u.encode() for u in urls]
    off = (np.arange(1025, dtype=np.uint64) * T)
    for tr in range(n_trunks):
        trunk = rng.integers(0, 50000, size=128, dtype=np.uint32)
        paths = np.empty((32, 32, T), np.uint32)
        paths[:, :, :128] = trunk
        paths[:, :, 128:256] = rng.integers(0, 50000, size=(32, 1, 128), dtype=np.uint32)
        paths[:, :, 256:] = rng.integers(0, 50000, size=(32, 32, 256), dtype=np.uint32)
        flat = np.ascontiguousarray(paths.reshape(-1))
        tens = (C.c_char_p * 1024)(*[url_c[(tr * 1024 + i) % W] for i in range(1024)])
        h.call("smgx_tree_insert_tokens_batch", model, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), 1024, tens)
        for i in rng.integers(0, 1024, size=max(1, 40000 // n_trunks)):
            sample.append(flat[i * T:(i + 1) * T].copy())
    t_build = time.time() - t0
    sample = np.stack(sample)

    def batch(seed):
        r = np.random.default_rng(seed)
        q = sample[r.integers(0, len(sample), size=B)].copy()
        u = r.random(B)
        for i in np.nonzero((u >= 0.8) & (u < 0.9))[0]:
            k = 16 * int(r.integers(1, 32))
            q[i, k:] = r.integers(0, 50000, size=T - k, dtype=np.uint32)
        nov = np.nonzero(u >= 0.9)[0]
        q[nov] = r.integers(0, 50000, size=(len(nov), T), dtype=np.uint32)
        return np.ascontiguousarray(q.reshape(-1))

    L = _lib.load()
    err = _lib.new_err()
    offs = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
    d_off = L.smgx_device_alloc(h.p, offs.nbytes, C.byref(err))
    h.call("smgx_memcpy_h2d", d_off, offs.ctypes.data_as(C.c_void_p), offs.nbytes)
    d_tok, d_out, d_info, infos = [], [], [], None
    for j in range(ring):
        q = batch(1000 + j)
        p = L.smgx_device_alloc(h.p, q.nbytes, C.byref(err)); h.call("smgx_memcpy_h2d", p, q.ctypes.data_as(C.c_void_p), q.nbytes)
        d_tok.append(p)
        d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
        d_info.append(L.smgx_device_alloc(h.p, B * 12, C.byref(err)))
    arr = lambda xs: (C.c_void_p * len(xs))(*xs)
    ns = (C.c_uint32 * per_call)(*([B] * per_call))

    def call(j0, want_info):
        js = [(j0 + k) % ring for k in range(per_call)]
        h.call("smgx_tree_walk_many_device", model, per_call, arr([d_tok[j] for j in js]), arr([d_off] * per_call), ns,
               arr([d_out[j] for j in js]), arr([d_info[j] for j in js]) if want_info else None)

    for w in range(3):
        call(w * per_call, True)
    h.call("smgx_synchronize")
    info = np.zeros(B, dtype=[("matched", "<u4"), ("input", "<u4"), ("branch", "u1"), ("nodes", "u1"), ("r", "u1", 2)])
    h.call("smgx_memcpy_d2h", info.ctypes.data_as(C.c_void_p), d_info[0], B * 12)
    # algorithmic bytes / decision: 4·T request tokens + 4·M label tokens compared + 96·N visited nodes
    alg = float(np.mean(4.0 * info["input"] + 4.0 * info["matched"] + 96.0 * info["nodes"]))
    h.call("smgx_timer_start_all")
    for s in range(steps):
        call(s * per_call, False)
    ms = C.c_float()
    h.call("smgx_timer_stop_all_ms", C.byref(ms))
    n = steps * per_call * B
    dps = n / (ms.value * 1e-3)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6486.8))
    return {"mode": "K2a token-tree walk + pick kernel only (read-only, HBM-resident requests; tree_select_kernel)",
            "workers": W, "batch": B, "tokens_per_request": T, "tree_nodes": n_trunks * (1 + 32 + 1024), "label_bytes": n_trunks * 266368 * 4,
            "tree_build_s": round(t_build, 1), "ring_batches": ring, "batches_per_call": per_call,
            "decisions_per_s": dps, "ms_per_batch": ms.value / (steps * per_call),
            "mix": {"matched_mean": float(info["matched"].mean()), "nodes_mean": float(info["nodes"].mean()),
                    "branches": {str(k): int(v) for k, v in zip(*np.unique(info["branch"], return_counts=True))}},
            "roofline": {"bound": "hbm", "alg_bytes_per_decision": alg, "achieved": alg * dps / 1e9, "peak": peak, "unit": "GB/s", "frac": alg * dps / 1e9 / peak}}


def text_mode():
    """HTTP text (string-tree) mode: chat-style routing texts with shared system prompts, ~2 KB each."""
    import random
    from oracle import orc
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy, synth
    W, B = 64, int(os.environ.get("BATCH", "1024"))
    bmode = os.environ.get("TREE_BATCH_MODE", "sequential")
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), max_batch=B, tree_batch_mode=bmode)
    ws = [BasicWorker(u) for u in urls]
    for w, l in zip(ws, synth.poisson_loads(W, 8, 42)):
        w.set_load(int(l))
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    op.set_state([w.load() for w in ws], [1] * W, [1] * W)
    r = random.Random(7)
    words = ["cache", "aware", "router", "prefix", "radix", "tree", "worker", "tenant", "load", "balance", "token", "你好", "été", "GPU"]
    systems = [" ".join(r.choice(words) for _ in range(200)) for _ in range(32)]      # ≈1.2 KB shared prefixes
    def batch():
        return [r.choice(systems) + " " + " ".join(r.choice(words) for _ in range(r.randrange(20, 160))) for _ in range(B)]
    batches = [batch() for _ in range(4)]
    snap = bmode == "snapshot"
    idx, _ = pol.select_worker_batch_request_text(ws, batches[0])
    want = op.select_batch_text(batches[0], snapshot=snap)[0]
    assert np.array_equal(idx, want), "text-mode picks differ from the oracle"
    from smg_b200.policy import TiktokenTokenizer
    enc = [TiktokenTokenizer._ragged(b) for b in batches[1:]]     # UTF-8 encoding of Python strings is not part of either timed region
    t0 = time.perf_counter()
    for data, offs in enc:
        pol.select_worker_batch_request_text(ws, want_info=False, data=data, offsets=offs)
    t_gpu = time.perf_counter() - t0
    t_cpu = sum(op.select_batch_text(b, snapshot=snap)[4] for b in batches[1:])
    assert pol.string_tree().entries() == op.string_tree().entries(), "trees diverged"
    n = 3 * B
    return {"mode": "HTTP text / string tree (cache_aware.rs:907-974): GPU walk+pick, host-side match effects + inserts in request order",
            "tree_batch_mode": bmode, "workers": W, "batch": B, "mean_request_bytes": int(np.mean([len(t.encode()) for t in batches[1]])),
            "smgx_decisions_per_s": n / t_gpu, "oracle_1core_decisions_per_s": n / t_cpu, "parity_first_batch": True,
            "trees_identical_after": True, "kernel_launches": pol.kernel_launches()}


def textwalk_mode():
    """K2c by itself: read-only walk + pick of HBM-resident text batches against a string tree of DOCS chat-shaped documents
    (32 shared ≈1.2 KB system prompts + distinct tails, ≈2 KB each)."""
    import random
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy, _lib, synth
    W, B = 64, 4096
    n_docs = int(os.environ.get("DOCS", "100000"))
    ring, steps, per_call = int(os.environ.get("RING", "32")), int(os.environ.get("STEPS", "20")), int(os.environ.get("PER_CALL", "8"))
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), max_batch=B)
    ws = [BasicWorker(u) for u in urls]
    for w, l in zip(ws, synth.poisson_loads(W, 8, 42)):
        w.set_load(int(l))
    pol.init_workers(ws)
    model = pol._push_fleet(ws)
    h, L = pol._h, _lib.load()
    r = random.Random(7)
    words = ["cache", "aware", "router", "prefix", "radix", "tree", "worker", "tenant", "load", "balance", "token", "你好", "été", "GPU", "latency", "batch"]
    systems = [" ".join(r.choice(words) for _ in range(200)) for _ in range(32)]
    def doc():
        return r.choice(systems) + " " + " ".join(r.choice(words) for _ in range(r.randrange(60, 160)))
    docs = [doc() for _ in range(n_docs)]
    tree = pol.string_tree()
    t0 = time.time()
    for i, d in enumerate(docs):
        tree.insert_text(d, urls[i % W])
    t_build = time.time() - t0

    def batch(seed):
        rr = random.Random(seed)
        out = []
        for _ in range(B):
            u = rr.random()
            d = docs[rr.randrange(n_docs)]
            if u < 0.8:
                out.append(d)
            elif u < 0.9:
                out.append(d[: rr.randrange(100, len(d))] + " novel tail " + " ".join(rr.choice(words) for _ in range(40)))
            else:
                out.append(" ".join(rr.choice(words[::-1]) for _ in range(300)))
        blobs = [t.encode() for t in out]
        offs = np.zeros(B + 1, np.uint32)
        np.cumsum([len(b) for b in blobs], out=offs[1:])
        return np.frombuffer(b"".join(blobs), np.uint8).copy(), offs

    err = _lib.new_err()
    d_txt, d_off, d_out, d_info, d_node = [], [], [], [], []
    total_bytes = 0
    for j in range(ring):
        data, offs = batch(500 + j)
        total_bytes += data.size
        pt = L.smgx_device_alloc(h.p, data.nbytes + 16, C.byref(err)); h.call("smgx_memcpy_h2d", pt, data.ctypes.data_as(C.c_void_p), data.nbytes)
        po = L.smgx_device_alloc(h.p, offs.nbytes, C.byref(err)); h.call("smgx_memcpy_h2d", po, offs.ctypes.data_as(C.c_void_p), offs.nbytes)
        d_txt.append(pt); d_off.append(po)
        d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
        d_info.append(L.smgx_device_alloc(h.p, B * 12, C.byref(err)))
        d_node.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
    arr = lambda xs: (C.c_void_p * len(xs))(*xs)
    ns = (C.c_uint32 * per_call)(*([B] * per_call))

    def call(j0, want_info):
        js = [(j0 + k) % ring for k in range(per_call)]
        h.call("smgx_stree_walk_many_device", model, per_call, arr([d_txt[j] for j in js]), arr([d_off[j] for j in js]), ns,
               arr([d_out[j] for j in js]), arr([d_info[j] for j in js]) if want_info else None, arr([d_node[j] for j in js]))

    for w in range(3):
        call(w * per_call, True)
    h.call("smgx_synchronize")
    info = np.zeros(B, dtype=[("matched", "<u4"), ("input", "<u4"), ("branch", "u1"), ("nodes", "u1"), ("r", "u1", 2)])
    h.call("smgx_memcpy_d2h", info.ctypes.data_as(C.c_void_p), d_info[0], B * 12)
    mean_bytes = total_bytes / (ring * B)
    # algorithmic bytes / decision: request bytes + label bytes compared (≈ matched chars × bytes/char) + 48 B per visited node
    bpc = mean_bytes / max(1.0, float(info["input"].mean()))
    alg = float(mean_bytes + info["matched"].mean() * bpc + 48.0 * info["nodes"].mean())
    h.call("smgx_timer_start_all")
    for s_ in range(steps):
        call(s_ * per_call, False)
    ms = C.c_float()
    h.call("smgx_timer_stop_all_ms", C.byref(ms))
    n = steps * per_call * B
    dps = n / (ms.value * 1e-3)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6486.8))
    return {"mode": "K2c string-tree walk + pick kernel only (read-only, HBM-resident text; string_select_kernel)", "workers": W, "batch": B,
            "tree_docs": n_docs, "tree_nodes": int(tree.node_count()), "tree_build_s": round(t_build, 1), "mean_request_bytes": mean_bytes,
            "decisions_per_s": dps, "ms_per_batch": ms.value / (steps * per_call),
            "mix": {"matched_chars_mean": float(info["matched"].mean()), "input_chars_mean": float(info["input"].mean()), "nodes_mean": float(info["nodes"].mean())},
            "roofline": {"bound": "hbm", "alg_bytes_per_decision": alg, "achieved": alg * dps / 1e9, "peak": peak, "unit": "GB/s", "frac": alg * dps / 1e9 / peak}}


def prefix_mode():
    """Adjacent policy: prefix_hash on the config-2 fleet shape — 64 workers, 512-token requests, batches of 4096,
    default PrefixHashConfig (first 256 tokens hashed).  Kernel-only (HBM-resident ring of batches larger than L2), through the C ABI
    with host buffers, and the oracle on one host core."""
    from oracle import orc
    from smg_b200 import BasicWorker, PrefixHashPolicy, _lib, synth
    W, B, T = 64, int(os.environ.get("BATCH", "4096")), 512
    ring_n, steps, per_call = int(os.environ.get("RING", "32")), int(os.environ.get("STEPS", "50")), int(os.environ.get("PER_CALL", "32"))
    urls = synth.worker_urls(W)
    pol = PrefixHashPolicy(max_batch=B)
    ws = [BasicWorker(u) for u in urls]
    loads = synth.poisson_loads(W, 8, 42)
    for w, l in zip(ws, loads):
        w.set_load(int(l))
    ring = pol.hash_ring(urls)
    model = pol._push_fleet(ws, ring)
    h = pol._h
    L = _lib.load()
    err = _lib.new_err()
    offs = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
    d_off = L.smgx_device_alloc(h.p, offs.nbytes, C.byref(err))
    h.call("smgx_memcpy_h2d", d_off, offs.ctypes.data_as(C.c_void_p), offs.nbytes)
    # shared system prompts: 256 distinct 256-token prefixes, Zipf-popular, each followed by a unique 256-token tail
    rng = np.random.default_rng(42)
    prefixes = rng.integers(0, 128000, size=(256, 256), dtype=np.uint32)
    zw = 1.0 / np.arange(1, 257) ** 1.1
    zw /= zw.sum()
    host, d_tok, d_out = [], [], []
    for j in range(ring_n):
        q = rng.integers(0, 128000, size=(B, T), dtype=np.uint32)
        shared = rng.random(B) < 0.7
        q[shared, :256] = prefixes[rng.choice(256, size=int(shared.sum()), p=zw)]
        q = np.ascontiguousarray(q.reshape(-1))
        host.append(q)
        ptr = L.smgx_device_alloc(h.p, q.nbytes, C.byref(err)); h.call("smgx_memcpy_h2d", ptr, q.ctypes.data_as(C.c_void_p), q.nbytes)
        d_tok.append(ptr)
        d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
    arr = lambda xs: (C.c_void_p * len(xs))(*xs)
    ns = (C.c_uint32 * per_call)(*([B] * per_call))

    def call(j0):
        js = [(j0 + k) % ring_n for k in range(per_call)]
        h.call("smgx_prefix_hash_select_many_tokens_device", model, per_call, arr([d_tok[j] for j in js]), arr([d_off] * per_call), ns, arr([d_out[j] for j in js]))

    for w in range(3):
        call(w * per_call)
    h.call("smgx_synchronize")
    l0 = pol.kernel_launches()
    h.call("smgx_timer_start_all")           # the pick kernels run on a side lane: time across all lanes
    for s in range(steps):
        call(s * per_call)
    ms = C.c_float()
    h.call("smgx_timer_stop_all_ms", C.byref(ms))
    launches = pol.kernel_launches() - l0
    n = steps * per_call * B
    dps = n / (ms.value * 1e-3)
    got = np.zeros(B, np.int32)
    h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[0], B * 4)
    # oracle: one core, same batch; also the parity check of this run
    opol, oring = orc.PrefixHashPolicy(), orc.HashRing(urls)
    oidx, obr, secs = opol.select_batch(urls, loads, [1] * W, oring, host[0], offs.astype(np.uint64))
    assert np.array_equal(got, oidx), "GPU picks differ from the oracle"
    cpu_secs = min(opol.select_batch(urls, loads, [1] * W, oring, host[j], offs.astype(np.uint64))[2] for j in range(1, 4))
    # through the C ABI with host buffers (pageable numpy arrays in, results out), synchronous calls
    e_steps = int(os.environ.get("E2E_STEPS", "200"))
    out = np.zeros(B, np.int32)
    for j in range(3):
        h.call("smgx_prefix_hash_select_batch_tokens", model, host[j].ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), B, None,
               out.ctypes.data_as(C.c_void_p), None)
    t0 = time.perf_counter()
    for s in range(e_steps):
        h.call("smgx_prefix_hash_select_batch_tokens", model, host[s % ring_n].ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), B, None,
               out.ctypes.data_as(C.c_void_p), None)
    e2e = e_steps * B / (time.perf_counter() - t0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6486.8))
    alg = 256 * 4 + 4 + 4   # hashed prefix + offset + pick
    return {"mode": "prefix_hash policy (prefix_hash_kernel + prefix_pick_kernel): XXH3 of the first 256 tokens → blake3 ring lookup → bounded-load pick",
            "workers": W, "batch": B, "tokens_per_request": T, "prefix_token_count": 256, "ring_entries": len(ring), "ring_batches": ring_n,
            "input_bytes_resident": ring_n * B * T * 4, "batches_per_launch": per_call, "launches": int(launches),
            "decisions_per_s": dps, "ms_per_batch": ms.value / (steps * per_call),
            "branches": {orc.PREFIX_BRANCHES[int(k)]: int(v) for k, v in zip(*np.unique(obr, return_counts=True))},
            "distinct_workers_picked": int(len(set(got.tolist()))),
            "e2e_host_buffers_decisions_per_s": e2e, "h2d_bytes_per_step": B * 256 * 4 + (B + 1) * 4, "d2h_bytes_per_step": B * 4,
            "cpu_baseline": {"value": B / cpu_secs, "unit": "decisions/s", "cores": 1, "kind": "port", "sample": "3 batches of 4096, best"},
            "roofline": {"bound": "hbm", "alg_bytes_per_decision": alg, "achieved": alg * dps / 1e9, "peak": peak, "unit": "GB/s", "frac": alg * dps / 1e9 / peak}}


def ingest_mode():
    """KV-event ingest: batches of 1024 Stored events (one 512-token sequence = 32 blocks of 16 tokens each, for worker
    i mod 64) through smgx_kv_events_apply — token_ids hashed in one GPU launch per batch, index writers on the host — vs the oracle's
    apply_event path on one core (hash on the CPU, then apply_stored)."""
    from oracle import orc
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy, _lib, synth
    W, E, P, bs = 64, 1024, 32, 16
    n_batches = int(os.environ.get("BATCHES", "16"))
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG))
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", 64)
    pol.set_kv_event_monitor(mon)
    oix = orc.PositionalIndexer(64)
    for u in synth.worker_urls(W):
        ix.intern_worker(u); oix.intern_worker(u)
    seqs = synth.gen_sequences(E * n_batches, P * bs, 7)
    Lo = orc.lib()
    t_gpu = t_cpu = 0.0
    offs = (np.arange(E * P + 1, dtype=np.uint64) * bs).astype(np.uint32)
    for b in range(n_batches):
        toks = np.ascontiguousarray(seqs[b * E:(b + 1) * E].reshape(-1))
        hashes = np.arange(1 + b * E * P, 1 + (b + 1) * E * P, dtype=np.int64)
        evs = (_lib.KvEvent * E)()
        for e in range(E):
            evs[e].kind, evs[e].worker_id, evs[e].first_block, evs[e].n_blocks, evs[e].has_parent = 0, (b * E + e) % W, e * P, P, 0
        fb = C.c_uint32()
        t0 = time.perf_counter()
        pol._h.call("smgx_kv_events_apply", b"unknown", C.cast(evs, C.c_void_p), E, hashes.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                    toks.ctypes.data_as(C.c_void_p), E * P, C.byref(fb))
        t_gpu += time.perf_counter() - t0
        content = np.zeros(E * P, np.uint64)
        useq = hashes.view(np.uint64)
        t0 = time.perf_counter()
        for e in range(E):
            row = toks[e * P * bs:(e + 1) * P * bs]
            Lo.orc_request_content_hashes(row.ctypes.data_as(C.c_void_p), P * bs, bs, content[e * P:].ctypes.data_as(C.c_void_p), P)
            Lo.orc_indexer_apply_stored(oix.h, (b * E + e) % W, useq[e * P:].ctypes.data_as(C.c_void_p), content[e * P:].ctypes.data_as(C.c_void_p), P, 0, 0)
        t_cpu += time.perf_counter() - t0
    assert ix.current_size() == oix.current_size() and ix.entry_count() == oix.entry_count()
    n_ev = E * n_batches
    return {"mode": "KV-event ingest: smgx_kv_events_apply (token_ids hashed on the GPU, one launch per 1024-event batch; index writers on the host mirror)",
            "events": n_ev, "blocks_per_event": P, "tokens_per_block": bs, "smgx_events_per_s": n_ev / t_gpu, "smgx_blocks_per_s": n_ev * P / t_gpu,
            "oracle_1core_events_per_s": n_ev / t_cpu, "index_entries_after": int(ix.entry_count()), "state_equal": True}


def sharded_mode():
    import torch
    import torch.distributed as dist
    from smg_b200 import CacheAwareConfig, _lib, synth
    from smg_b200.sharding import CAND_BYTES, FLEET_BYTES, ShardedEventRouter
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    W = 512 * world                    # config 4: 512 workers per GPU
    B, T, bs = 4096, 512, 16
    urls = synth.worker_urls(W)
    router = ShardedEventRouter(urls, rank, world, CacheAwareConfig(eviction_interval_secs=0, **CFG), jump_size=64, device_id=local,
                                max_batch=B, max_tokens_per_request=T)
    n_seq = 31250
    seqs = synth.gen_sequences(n_seq, T, 44)
    flat = np.ascontiguousarray(seqs.reshape(-1))
    P = T // bs
    hashes = np.zeros(n_seq * P, np.uint64)
    got = C.c_uint32()
    for off in range(0, flat.size, 4096 * T):
        part = flat[off:off + 4096 * T]
        router.policy._h.call("smgx_content_hashes", part.ctypes.data_as(C.c_void_p), part.size, bs, hashes[off // bs:].ctypes.data_as(C.c_void_p),
                              part.size // bs, C.byref(got))
    ids = np.arange(1, n_seq * P + 1, dtype=np.uint64)
    for s in range(n_seq):
        g = s % W
        if router.owns(g):
            router.policy._h.call("smgx_indexer_apply_stored", router.indexer.model, router.local_id(g), ids[s * P:(s + 1) * P].ctypes.data_as(C.c_void_p),
                                  hashes[s * P:(s + 1) * P].ctypes.data_as(C.c_void_p), P, None)
    router.set_fleet_state(synth.poisson_loads(W, 8, 44), np.ones(W, np.uint8))
    h, L = router._h, _lib.load()
    q = synth.gen_queries(seqs, B, 44)
    tokens, offsets = synth.ragged(q)
    d_tok = torch.from_numpy(tokens.view(np.int32)).cuda()
    d_off = torch.from_numpy(offsets.view(np.int32)).cuda()
    cand = torch.empty(B * CAND_BYTES, dtype=torch.uint8, device="cuda")
    fleet = torch.empty(FLEET_BYTES, dtype=torch.uint8, device="cuda")
    all_c = torch.empty(world * B * CAND_BYTES, dtype=torch.uint8, device="cuda")
    all_f = torch.empty(world * FLEET_BYTES, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, dtype=torch.int32, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())

    fused = os.environ.get("FUSED", "1") != "0"
    if fused:
        def ag(arr):
            t = torch.from_numpy(arr.copy()).cuda()
            out = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(out, t)
            return out.cpu().numpy()
        router.connect_peers(ag)

    def step_fused():
        router.select_fused_device(vp(d_tok), vp(d_off), B, T, vp(out))      # asynchronous; stream order does the rest

    def step():
        if fused:
            return step_fused()
        h.call("smgx_shard_candidates_device", router.model, 0, vp(d_tok), vp(d_off), B, T, vp(cand), vp(fleet))
        h.call("smgx_synchronize")                    # the library's lane stream → torch's stream
        dist.all_gather_into_tensor(all_c, cand)
        dist.all_gather_into_tensor(all_f, fleet)
        torch.cuda.synchronize()
        h.call("smgx_shard_reduce_device", 0, vp(all_c), vp(all_f), router.gbase.ctypes.data_as(C.c_void_p), world, B, vp(out), None)
        h.call("smgx_synchronize")

    for _ in range(5):
        step()
    h.call("smgx_synchronize")
    ref_picks = None
    if fused:   # same picks as the collective path
        fused = False
        step()
        ref_picks = out.cpu().numpy().copy()
        fused = True
        step()
        h.call("smgx_synchronize")
        assert np.array_equal(out.cpu().numpy(), ref_picks), "fused exchange and all-gather path disagree"
    dist.barrier()
    torch.cuda.synchronize()
    K = int(os.environ.get("STEPS", "200"))
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    h.call("smgx_synchronize")
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    picks = out.cpu().numpy()
    res = None
    if rank == 0:
        res = {"mode": "worker-id-sharded event pick (config 4 shape): candidates kernels per shard → " +
                       ("peer-memory exchange (stores into every rank's gather buffer over NVLink + flags) → merge kernel, one stream, no host sync"
                        if fused else "NCCL all-gather (24 B/request/shard) → merge kernel"),
               "exchange": "peer-memory" if fused else "nccl",
               "n_gpus": world, "workers": W, "workers_per_gpu": 512, "batch": B, "index_entries_total": n_seq * P,
               "decisions_per_s": K * B / float(dt.item()), "ms_per_step": 1e3 * float(dt.item()) / K,
               "picks_in_range": bool(picks.min() >= 0 and picks.max() < W), "distinct_workers_picked": int(len(set(picks.tolist())))}
    dist.barrier()
    dist.destroy_process_group()
    return res


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "tree"
    r = {"tree": tree_mode, "treewalk": treewalk_mode, "text": text_mode, "textwalk": textwalk_mode, "ingest": ingest_mode, "sharded": sharded_mode, "prefix": prefix_mode}[mode]()
    if r is not None:
        print(json.dumps(r), flush=True)
