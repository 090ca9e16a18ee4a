"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Mirrors the reference's Rust API names so the ported unit tests read like the originals
(crates/kv_index/src/{event_tree,token_tree,string_tree}.rs, model_gateway/src/policies/cache_aware.rs).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
        vp, sz, u64, u32, i64, cp = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int64, C.c_char_p
        P = C.POINTER

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)

        sig("orc_reset_globals", None)
        sig("orc_token_ts", u64)
        sig("orc_string_epoch", u64)
        sig("orc_xxh3_64", u64, vp, sz, u64)
        sig("orc_content_hash", u64, vp, sz)
        sig("orc_next_seq_hash", u64, u64, u64)
        sig("orc_request_content_hashes", sz, vp, sz, sz, vp, sz)
        sig("orc_indexer_new", vp, sz)
        sig("orc_indexer_free", None, vp)
        sig("orc_indexer_intern_worker", u32, vp, cp)
        sig("orc_indexer_worker_id", i64, vp, cp)
        sig("orc_indexer_apply_stored", C.c_int, vp, u32, vp, vp, sz, C.c_int, u64)
        sig("orc_indexer_apply_removed", None, vp, u32, vp, sz)
        sig("orc_indexer_apply_cleared", None, vp, u32)
        sig("orc_indexer_remove_worker", None, vp, u32)
        sig("orc_indexer_current_size", sz, vp)
        sig("orc_indexer_entry_count", sz, vp)
        sig("orc_indexer_tree_size", sz, vp, u32)
        sig("orc_indexer_find_matches", sz, vp, vp, sz, C.c_int, vp, vp, vp, sz)
        sig("orc_ttree_new", vp, C.c_int)
        sig("orc_ttree_free", None, vp)
        sig("orc_ttree_insert", None, vp, vp, sz, cp)
        sig("orc_ttree_match", sz, vp, vp, sz, vp, sz, vp, vp, sz)
        sig("orc_ttree_evict_tenant", None, vp, cp, sz)
        sig("orc_ttree_evict_by_size", None, vp, sz)
        sig("orc_ttree_tenant_size", sz, vp, cp)
        sig("orc_ttree_node_count", sz, vp)
        sig("orc_ttree_clear", None, vp)
        sig("orc_ttree_set_priority", None, vp, vp, sz, C.c_int32)
        sig("orc_ttree_entries", sz, vp, vp, sz)
        sig("orc_stree_new", vp)
        sig("orc_stree_free", None, vp)
        sig("orc_stree_insert", None, vp, cp, sz, cp)
        sig("orc_stree_match", sz, vp, cp, sz, vp, sz, vp, vp, sz)
        sig("orc_stree_prefix_match_tenant", sz, vp, cp, sz, cp, vp, sz)
        sig("orc_stree_force_cached_tenant", C.c_int, vp, cp, sz, cp)
        sig("orc_stree_evict_by_size", None, vp, sz)
        sig("orc_stree_evict_by_tenant", None, vp, cp, sz)
        sig("orc_stree_remove_tenant_all", None, vp, cp)
        sig("orc_stree_tenant_size", sz, vp, cp)
        sig("orc_stree_node_count", sz, vp)
        sig("orc_stree_used_sizes", sz, vp, vp, sz)
        sig("orc_stree_char_counts", sz, vp, vp, sz)
        sig("orc_stree_entries", sz, vp, vp, sz)
        sig("orc_stree_snapshot_bytes", sz, vp, vp, sz)
        sig("orc_stree_load_snapshot_bytes", C.c_int, vp, vp, sz)
        sig("orc_stree_merge_snapshot_bytes", C.c_int, vp, vp, sz)
        sig("orc_policy_new", vp, C.c_float, u64, C.c_float, u64, u64, u64)
        sig("orc_policy_free", None, vp)
        sig("orc_policy_set_workers", None, vp, P(cp), P(cp), sz, C.c_int)
        sig("orc_policy_set_state", None, vp, vp, vp, vp, sz)
        sig("orc_policy_processed", u64, vp, sz)
        sig("orc_policy_set_monitor", None, vp, C.c_int)
        sig("orc_policy_attach_indexer", None, vp, cp, vp)
        sig("orc_policy_set_block_size", None, vp, cp, sz)
        sig("orc_policy_has_event_indexer", C.c_int, vp, cp)
        sig("orc_policy_evict_cache", None, vp, sz)
        sig("orc_policy_token_tree", vp, vp, cp)
        sig("orc_policy_string_tree", vp, vp, cp)
        sig("orc_policy_select", None, vp, cp, sz, C.c_int, vp, sz, C.c_int, vp, vp, sz)
        sig("orc_policy_select_batch_tokens", C.c_double, vp, vp, vp, sz, vp, vp, vp)
        sig("orc_policy_select_batch_tokens_snapshot", C.c_double, vp, vp, vp, sz, vp, vp, vp)
        sig("orc_policy_select_batch_text", C.c_double, vp, vp, vp, sz, C.c_int, vp, vp, vp, vp)
        sig("orc_blake3", None, vp, sz, vp)
        sig("orc_hash_path_bytes", C.c_uint64, vp, sz)
        sig("orc_hash_token_path", C.c_uint64, vp, sz)
        sig("orc_policy_hash_index", sz, vp, cp, C.c_int, vp, sz)
        sig("orc_policy_select_steps_mt", C.c_double, vp, vp, vp, sz, sz, sz, vp, C.c_int, C.c_int)
        sig("orc_policy_select_batch_tokens_feedback", C.c_double, vp, vp, vp, sz, vp, vp, vp, vp)
        sig("orc_tuned_select_steps_mt", C.c_double, vp, vp, C.c_float, C.c_uint64, sz, vp, vp, sz, sz, sz, vp, C.c_int)
        sig("orc_policy_apply_known_remote_insert", C.c_int, vp, cp, C.c_int, u64, cp)
        sig("orc_policy_apply_repair_entry", None, vp, cp, C.c_int, vp, sz, cp)
        sig("orc_ring_new", vp, P(cp), sz)
        sig("orc_ring_free", None, vp)
        sig("orc_ring_len", sz, vp)
        sig("orc_ring_worker_count", sz, vp)
        sig("orc_ring_hash_position", u64, cp, sz)
        sig("orc_ring_entries", sz, vp, vp, vp, sz)
        sig("orc_ring_find_healthy", i64, vp, cp, sz, vp)
        sig("orc_prefix_hash", u64, vp, sz, sz)
        sig("orc_prefix_load_ok", C.c_int, C.c_double, u64, u64, sz)
        sig("orc_prefix_select_batch", C.c_double, sz, C.c_double, P(cp), vp, vp, sz, vp, vp, vp, sz, C.c_int, vp, vp)
        sig("orc_p2c_select_batch", None, P(cp), vp, vp, vp, sz, P(cp), vp, sz, C.c_uint64, sz, vp, vp, vp, vp)
        sig("orc_p2c_effective_token_usage", C.c_double, vp, sz)
        _lib = L
    return _lib


def _u32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


def reset_globals():
    lib().orc_reset_globals()


def xxh3_64(data: bytes, seed: int = 0) -> int:
    buf = C.create_string_buffer(data, len(data))
    return lib().orc_xxh3_64(C.cast(buf, C.c_void_p), len(data), seed)


def blake3_digest(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_blake3(C.c_char_p(data), len(data), C.cast(out, C.c_void_p))
    return out.raw


def hash_node_path(text: str) -> int:       # crates/mesh/src/hash.rs:22-30
    b = text.encode("utf-8")
    return lib().orc_hash_path_bytes(C.c_char_p(b), len(b))


def hash_token_path(tokens) -> int:         # crates/mesh/src/hash.rs:40-52
    t = _u32(tokens)
    return lib().orc_hash_token_path(_ptr(t), t.size)


def compute_content_hash(tokens) -> int:
    t = _u32(tokens)
    return lib().orc_content_hash(_ptr(t), t.size)


def compute_next_seq_hash(prev: int, cur: int) -> int:
    return lib().orc_next_seq_hash(prev, cur)


def compute_request_content_hashes(tokens, block_size):
    t = _u32(tokens)
    cap = (t.size // block_size) if block_size else 0
    out = np.zeros(max(cap, 1), dtype=np.uint64)
    n = lib().orc_request_content_hashes(_ptr(t), t.size, block_size, _ptr(out), cap)
    return [int(x) for x in out[:n]]


WORKER_NOT_TRACKED = "WorkerNotTracked"
PARENT_BLOCK_NOT_FOUND = "ParentBlockNotFound"


class ApplyError(Exception):
    pass


class PositionalIndexer:
    """kv_index::PositionalIndexer.  Blocks are (seq_hash, content_hash) pairs; the per-worker
    WorkerBlockMap the reference makes caller-owned lives inside the handle, keyed by worker id."""

    def __init__(self, jump_size=32):
        if jump_size <= 0:
            raise ValueError("jump_size must be greater than 0")
        self.h = lib().orc_indexer_new(jump_size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_indexer_free(self.h)
            self.h = None

    def intern_worker(self, url):
        return lib().orc_indexer_intern_worker(self.h, url.encode())

    def worker_id(self, url):
        r = lib().orc_indexer_worker_id(self.h, url.encode())
        return None if r < 0 else r

    def apply_stored(self, wid, blocks, parent=None):
        seq = _u64([b[0] for b in blocks])
        con = _u64([b[1] for b in blocks])
        rc = lib().orc_indexer_apply_stored(self.h, wid, _ptr(seq), _ptr(con), len(blocks), 0 if parent is None else 1,
                                            0 if parent is None else parent)
        if rc == 1:
            raise ApplyError(WORKER_NOT_TRACKED)
        if rc == 2:
            raise ApplyError(PARENT_BLOCK_NOT_FOUND)

    def apply_removed(self, wid, seq_hashes):
        s = _u64(seq_hashes)
        lib().orc_indexer_apply_removed(self.h, wid, _ptr(s), s.size)

    def apply_cleared(self, wid):
        lib().orc_indexer_apply_cleared(self.h, wid)

    def remove_worker(self, wid):
        lib().orc_indexer_remove_worker(self.h, wid)

    def current_size(self):
        return lib().orc_indexer_current_size(self.h)

    def entry_count(self):
        return lib().orc_indexer_entry_count(self.h)

    def find_matches(self, content_hashes, early_exit=False):
        hs = _u64(content_hashes)
        cap = 4096
        ids = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.uint32)
        ts = np.zeros(cap, np.uint64)
        n = lib().orc_indexer_find_matches(self.h, _ptr(hs), hs.size, 1 if early_exit else 0, _ptr(ids), _ptr(sc), _ptr(ts), cap)
        assert n <= cap
        return ({int(ids[i]): int(sc[i]) for i in range(n)}, {int(ids[i]): int(ts[i]) for i in range(n)})


def apply_kv_events(indexer: "PositionalIndexer", worker_id: int, events):
    """KvEventMonitor::apply_event over one batch (model_gateway/src/worker/kv_event_monitor.rs:525-597): Stored → convert_kv_block
    (content hash of the block's own token_ids, i64 block hashes reinterpreted as u64 :592-597) + apply_stored with the fresh-chain
    retry on ApplyError (:559-571); Removed → apply_removed; Cleared → apply_cleared.  Returns the number of fallbacks."""
    fallbacks = 0
    for ev in events:
        if "stored" in ev:
            st = ev["stored"]
            blocks = [(int(b["block_hash"]) & 0xFFFFFFFFFFFFFFFF, compute_content_hash(b.get("token_ids", []))) for b in st.get("blocks", [])]
            parent = st.get("parent_block_hash")
            try:
                indexer.apply_stored(worker_id, blocks, None if parent is None else int(parent) & 0xFFFFFFFFFFFFFFFF)
            except ApplyError:
                fallbacks += 1
                try:
                    indexer.apply_stored(worker_id, blocks, None)
                except ApplyError:
                    pass
        elif "removed" in ev:
            indexer.apply_removed(worker_id, [int(h) & 0xFFFFFFFFFFFFFFFF for h in ev["removed"].get("block_hashes", [])])
        elif "cleared" in ev:
            indexer.apply_cleared(worker_id)
    return fallbacks


class TokenMatch:
    def __init__(self, tenant, matched, inp, nodes, edge, valid):
        self.tenant, self.matched_token_count, self.input_token_count = tenant, matched, inp
        self.nodes_visited, self.edge_tokens_compared, self.valid = nodes, edge, valid


LRU, LFU, FIFO, MRU, FILO, PRIORITY = range(6)


class TokenTree:
    def __init__(self, policy=LRU, handle=None):
        self.owned = handle is None
        self.h = lib().orc_ttree_new(policy) if handle is None else handle

    def __del__(self):
        if getattr(self, "h", None) and self.owned:
            lib().orc_ttree_free(self.h)
            self.h = None

    def insert_tokens(self, tokens, tenant):
        t = _u32(tokens)
        lib().orc_ttree_insert(self.h, _ptr(t), t.size, tenant.encode())

    def match_prefix_with_counts(self, tokens):
        t = _u32(tokens)
        ten = C.create_string_buffer(1024)
        val = C.create_string_buffer(1 << 16)
        cnt = np.zeros(4, np.uint64)
        lib().orc_ttree_match(self.h, _ptr(t), t.size, C.cast(ten, C.c_void_p), 1024, _ptr(cnt), C.cast(val, C.c_void_p), 1 << 16)
        valid = [v for v in val.value.decode().split("\n") if v]
        return TokenMatch(ten.value.decode(), int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3]), valid)

    def prefix_match_legacy(self, tokens):
        r = self.match_prefix_with_counts(tokens)
        return list(tokens[: r.matched_token_count]), r.tenant

    def evict_tenant(self, tenant, max_tokens):
        lib().orc_ttree_evict_tenant(self.h, tenant.encode(), max_tokens)

    def evict_tenant_by_size(self, max_size):
        lib().orc_ttree_evict_by_size(self.h, max_size)

    def tenant_token_size(self, tenant):
        return lib().orc_ttree_tenant_size(self.h, tenant.encode())

    def node_count(self):
        return lib().orc_ttree_node_count(self.h)

    def clear(self):
        lib().orc_ttree_clear(self.h)

    def set_priority(self, tokens, p):
        t = _u32(tokens)
        lib().orc_ttree_set_priority(self.h, _ptr(t), t.size, p)

    def entries(self):
        n = lib().orc_ttree_entries(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().orc_ttree_entries(self.h, C.cast(buf, C.c_void_p), n + 1)
        out = []
        for line in buf.value.decode().split("\n"):
            if not line:
                continue
            toks, tens = line.split("|")
            out.append(([int(x) for x in toks.split(",")] if toks else [],
                        [(kv.rsplit("=", 1)[0], int(kv.rsplit("=", 1)[1])) for kv in tens.split(";") if kv]))
        return out


class StringMatch:
    def __init__(self, tenant, matched, inp, nodes, valid):
        self.tenant, self.matched_char_count, self.input_char_count, self.nodes_visited, self.valid = tenant, matched, inp, nodes, valid


class Tree:
    """kv_index::Tree (string tree)."""

    def __init__(self, handle=None):
        self.owned = handle is None
        self.h = lib().orc_stree_new() if handle is None else handle

    def __del__(self):
        if getattr(self, "h", None) and self.owned:
            lib().orc_stree_free(self.h)
            self.h = None

    def insert_text(self, text, tenant):
        b = text.encode()
        lib().orc_stree_insert(self.h, b, len(b), tenant.encode())

    def match_prefix_with_counts(self, text):
        b = text.encode()
        ten = C.create_string_buffer(1024)
        val = C.create_string_buffer(1 << 16)
        cnt = np.zeros(4, np.uint64)
        lib().orc_stree_match(self.h, b, len(b), C.cast(ten, C.c_void_p), 1024, _ptr(cnt), C.cast(val, C.c_void_p), 1 << 16)
        valid = [v for v in val.value.decode().split("\n") if v]
        return StringMatch(ten.value.decode(), int(cnt[0]), int(cnt[1]), int(cnt[2]), valid)

    def prefix_match_legacy(self, text):
        r = self.match_prefix_with_counts(text)
        return text[: r.matched_char_count], r.tenant

    def prefix_match_tenant(self, text, tenant):
        b = text.encode()
        buf = C.create_string_buffer(len(b) + 1)   # one call: it draws an epoch (string_tree.rs:713-717)
        n = lib().orc_stree_prefix_match_tenant(self.h, b, len(b), tenant.encode(), C.cast(buf, C.c_void_p), len(b) + 1)
        return buf.raw[:n].decode()

    def force_cached_tenant(self, text, tenant):
        b = text.encode()
        return bool(lib().orc_stree_force_cached_tenant(self.h, b, len(b), tenant.encode()))

    def evict_tenant_by_size(self, max_size):
        lib().orc_stree_evict_by_size(self.h, max_size)

    def evict_by_tenant(self, tenant, max_chars):
        lib().orc_stree_evict_by_tenant(self.h, tenant.encode(), max_chars)

    def remove_tenant_all(self, tenant):
        lib().orc_stree_remove_tenant_all(self.h, tenant.encode())

    def tenant_char_size(self, tenant):
        return lib().orc_stree_tenant_size(self.h, tenant.encode())

    def node_count(self):
        return lib().orc_stree_node_count(self.h)

    def _kv(self, fn):
        n = fn(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        fn(self.h, C.cast(buf, C.c_void_p), n + 1)
        out = {}
        for line in buf.value.decode().split("\n"):
            if line:
                k, v = line.rsplit("=", 1)
                out[k] = int(v)
        return out

    def get_used_size_per_tenant(self):
        return self._kv(lib().orc_stree_used_sizes)

    def get_tenant_char_count(self):
        return self._kv(lib().orc_stree_char_counts)

    def entries(self):
        n = lib().orc_stree_entries(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().orc_stree_entries(self.h, C.cast(buf, C.c_void_p), n + 1)
        out = []
        for rec in buf.raw[:n].decode().split("\x1e"):
            if not rec:
                continue
            path, tens = rec.split("\x1f")
            out.append((path, [(kv.rsplit("=", 1)[0], int(kv.rsplit("=", 1)[1])) for kv in tens.split(";") if kv]))
        return out

    # ---- mesh wire format: kv_index::snapshot::TreeSnapshot (snapshot.rs) ----
    def snapshot_bytes(self) -> bytes:       # Tree::snapshot().to_bytes() (string_tree.rs:1066, snapshot.rs:44)
        n = lib().orc_stree_snapshot_bytes(self.h, None, 0)
        buf = C.create_string_buffer(max(n, 1))
        lib().orc_stree_snapshot_bytes(self.h, C.cast(buf, C.c_void_p), n)
        return buf.raw[:n]

    @classmethod
    def from_snapshot_bytes(cls, data: bytes):   # TreeSnapshot::from_bytes + Tree::from_snapshot (:1228)
        t = cls()
        if not lib().orc_stree_load_snapshot_bytes(t.h, data, len(data)):
            raise ValueError("bincode: malformed TreeSnapshot")
        return t

    def merge_snapshot_bytes(self, data: bytes):   # Tree::merge_snapshot (:1318)
        if not lib().orc_stree_merge_snapshot_bytes(self.h, data, len(data)):
            raise ValueError("bincode: malformed TreeSnapshot")


def decode_snapshot(data: bytes, allow_trailing: bool = False):
    """bincode 1.3 (default options) TreeSnapshot → [(edge, [(tenant, epoch)], child_count)], in plain Python (an independent reader
    of the wire format for the tests)."""
    import struct
    at, out = 0, []
    (n,) = struct.unpack_from("<Q", data, at); at += 8
    for _ in range(n):
        (ln,) = struct.unpack_from("<Q", data, at); at += 8
        edge = data[at:at + ln].decode("utf-8"); at += ln
        (nt,) = struct.unpack_from("<Q", data, at); at += 8
        tens = []
        for _ in range(nt):
            (tl,) = struct.unpack_from("<Q", data, at); at += 8
            name = data[at:at + tl].decode("utf-8"); at += tl
            (ep,) = struct.unpack_from("<Q", data, at); at += 8
            tens.append((name, ep))
        (cc,) = struct.unpack_from("<I", data, at); at += 4
        out.append((edge, tens, cc))
    assert allow_trailing or at == len(data)
    return out


def encode_snapshot(nodes) -> bytes:
    import struct
    b = struct.pack("<Q", len(nodes))
    for edge, tens, cc in nodes:
        e = edge.encode("utf-8")
        b += struct.pack("<Q", len(e)) + e + struct.pack("<Q", len(tens))
        for name, ep in tens:
            t = name.encode("utf-8")
            b += struct.pack("<Q", len(t)) + t + struct.pack("<Q", ep)
        b += struct.pack("<I", cc)
    return b


BRANCHES = ["no_healthy", "imbalanced_min_load", "event_overlap", "event_min_load", "tree_match", "tree_min_load",
            "tree_fallback_first_healthy", "no_tree_random"]


class Decision:
    def __init__(self, out, valid):
        self.idx = None if out[0] < 0 else int(out[0])
        self.branch = BRANCHES[int(out[1])]
        self.matched, self.input, self.score = int(out[2]), int(out[3]), int(out[4])
        self.valid = valid


class CacheAwarePolicy:
    """policies::CacheAwarePolicy + the Worker scalars it reads.  Worker state is set by index into the slice."""

    def __init__(self, cache_threshold=0.5, balance_abs_threshold=32, balance_rel_threshold=1.1, eviction_interval_secs=0,
                 max_tree_size=10000, block_size=16):
        self.h = lib().orc_policy_new(cache_threshold, balance_abs_threshold, balance_rel_threshold, eviction_interval_secs,
                                      max_tree_size, block_size)
        self.n = 0
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_policy_free(self.h)
            self.h = None

    def set_workers(self, urls, models=None, init=True):
        self.n = len(urls)
        U = (C.c_char_p * len(urls))(*[u.encode() for u in urls])
        M = (C.c_char_p * len(urls))(*[(m or "").encode() for m in (models or [""] * len(urls))])
        lib().orc_policy_set_workers(self.h, U, M, len(urls), 1 if init else 0)

    def set_state(self, loads=None, healthy=None, circuit_ok=None):
        lo = _u64(loads) if loads is not None else None
        he = np.ascontiguousarray(np.asarray(healthy, dtype=np.uint8)) if healthy is not None else None
        ci = np.ascontiguousarray(np.asarray(circuit_ok, dtype=np.uint8)) if circuit_ok is not None else None
        lib().orc_policy_set_state(self.h, _ptr(lo) if lo is not None else None, _ptr(he) if he is not None else None,
                                   _ptr(ci) if ci is not None else None, self.n)

    def processed(self, i):
        return lib().orc_policy_processed(self.h, i)

    def set_kv_event_monitor(self, present=True):
        lib().orc_policy_set_monitor(self.h, 1 if present else 0)

    def attach_indexer(self, model, indexer):
        self._keep.append(indexer)
        lib().orc_policy_attach_indexer(self.h, model.encode(), indexer.h)

    def set_block_size(self, model, bs):
        lib().orc_policy_set_block_size(self.h, model.encode(), bs)

    def has_event_indexer(self, model):
        return bool(lib().orc_policy_has_event_indexer(self.h, model.encode()))

    def evict_cache(self, max_size):
        lib().orc_policy_evict_cache(self.h, max_size)

    def token_tree(self, model="unknown"):
        h = lib().orc_policy_token_tree(self.h, model.encode())
        return TokenTree(handle=h) if h else None

    def string_tree(self, model="unknown"):
        h = lib().orc_policy_string_tree(self.h, model.encode())
        return Tree(handle=h) if h else None

    def select_worker(self, request_text=None, tokens=None):
        out = np.zeros(6, np.int64)
        valid = np.zeros(max(self.n, 1) + 8, np.int64)
        tb = request_text.encode() if request_text is not None else b""
        tk = _u32(tokens) if tokens is not None else _u32([])
        lib().orc_policy_select(self.h, tb, len(tb), 1 if request_text is not None else 0, _ptr(tk), tk.size,
                                1 if tokens is not None else 0, _ptr(out), _ptr(valid), valid.size)
        return Decision(out, [int(v) for v in valid[: int(out[5])]])

    def select_steps_mt(self, batches, steps, threads, step_barrier=False):
        """`steps` batches routed back to back by `threads` persistent host threads (event mode, read-only index).
        batches: list of (tokens u32, offsets u64[n+1]) with equal n.  → (picks of the last step, seconds)."""
        toks = [_u32(b[0]) for b in batches]
        offs = [_u64(b[1]) for b in batches]
        n = offs[0].size - 1
        TP = (C.c_void_p * len(batches))(*[t.ctypes.data for t in toks])
        OP = (C.c_void_p * len(batches))(*[o.ctypes.data for o in offs])
        idx = np.zeros(n, np.int32)
        secs = lib().orc_policy_select_steps_mt(self.h, TP, OP, len(batches), n, steps, _ptr(idx), threads, 1 if step_barrier else 0)
        return idx, secs

    def select_batch_tokens_feedback(self, tokens, offsets):
        """The request stream with the router's WorkerLoadGuard (router.rs:319-321): each pick bumps its worker's load before the next
        request is routed.  → (idx, branch, matched, loads after the batch)."""
        tk, off = _u32(tokens), _u64(offsets)
        n = off.size - 1
        idx, br, ma = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.uint32)
        loads = np.zeros(max(self.n, 1), np.uint64)
        lib().orc_policy_select_batch_tokens_feedback(self.h, _ptr(tk), _ptr(off), n, _ptr(idx), _ptr(br), _ptr(ma), _ptr(loads))
        return idx, br, ma, loads[: self.n]

    def tuned_select_steps_mt(self, indexer, batches, steps, threads, rel_thr, abs_thr, block_size):
        """The "port-tuned" variant (oracle/tuned_event.h): same event-mode decisions from a flat bitset index built from `indexer`,
        worker threads created before the clock starts.  → (picks of the last step, seconds)."""
        toks = [_u32(b[0]) for b in batches]
        offs = [_u64(b[1]) for b in batches]
        n = offs[0].size - 1
        TP = (C.c_void_p * len(batches))(*[t.ctypes.data for t in toks])
        OP = (C.c_void_p * len(batches))(*[o.ctypes.data for o in offs])
        idx = np.zeros(n, np.int32)
        secs = lib().orc_tuned_select_steps_mt(self.h, indexer.h, float(rel_thr), int(abs_thr), int(block_size), TP, OP, len(batches), n, steps, _ptr(idx), threads)
        return idx, secs

    # ---- TreeHandle (cache_aware.rs:443-645) ----
    def apply_known_remote_insert(self, model_id, tree_kind, node_hash, worker_url) -> bool:
        """tree_kind: "string" | "token".  True iff the hash resolves locally; the stored prefix then gains worker_url as a tenant."""
        return bool(lib().orc_policy_apply_known_remote_insert(self.h, model_id.encode(), 1 if tree_kind == "token" else 0, int(node_hash), worker_url.encode()))

    def apply_repair_page(self, model_id, tree_kind, entries) -> int:
        """entries: [("string", path, [(tenant, epoch)]) | ("token", tokens, [(tenant, epoch)])].  Entries whose variant differs from the
        page kind are skipped (:606-613, :633-640).  Returns the number applied."""
        applied = 0
        for kind, path, tenants in entries:
            if kind != tree_kind:
                continue
            names = "\n".join(t for t, _ in tenants).encode()
            if kind == "token":
                t = _u32(path)
                lib().orc_policy_apply_repair_entry(self.h, model_id.encode(), 1, _ptr(t), t.size, names)
            else:
                b = path.encode("utf-8")
                lib().orc_policy_apply_repair_entry(self.h, model_id.encode(), 0, b, len(b), names)
            applied += 1
        return applied

    def hash_index(self, kind="tokens", model="unknown"):
        """hash_index[model].token_tree / .string_tree (cache_aware.rs:95-101): {path hash: matched prefix}."""
        tk = 1 if kind == "text" else 0
        n = lib().orc_policy_hash_index(self.h, model.encode(), tk, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().orc_policy_hash_index(self.h, model.encode(), tk, C.cast(buf, C.c_void_p), n + 1)
        out = {}
        for rec in buf.raw[:n].decode("utf-8").split("\x1e"):
            if rec:
                k, v = rec.split("=", 1)
                out[int(k)] = v if tk else [int(x) for x in v.split(",") if x]
        return out

    def select_batch_text(self, texts, snapshot=False):
        """texts: list of str.  Returns (idx, branch, matched_chars, input_chars, secs)."""
        enc = [t.encode("utf-8") for t in texts]
        blob = b"".join(enc)
        off = np.zeros(len(enc) + 1, np.uint64)
        np.cumsum([len(e) for e in enc], out=off[1:])
        n = len(enc)
        idx = np.zeros(n, np.int32); br = np.zeros(n, np.uint8); ma = np.zeros(n, np.uint32); inp = np.zeros(n, np.uint32)
        buf = C.create_string_buffer(blob, len(blob) + 1)
        secs = lib().orc_policy_select_batch_text(self.h, buf, _ptr(off), n, 1 if snapshot else 0, _ptr(idx), _ptr(br), _ptr(ma), _ptr(inp))
        return idx, br, ma, inp, secs

    def select_batch_tokens(self, tokens, offsets, threads=0, snapshot=False):
        tk = _u32(tokens)
        off = _u64(offsets)
        n = off.size - 1
        idx = np.zeros(n, np.int32)
        if threads and threads > 1:
            i2, secs = self.select_steps_mt([(tk, off)], 1, threads)
            return i2, None, None, secs
        br = np.zeros(n, np.uint8)
        ma = np.zeros(n, np.uint32)
        fn = lib().orc_policy_select_batch_tokens_snapshot if snapshot else lib().orc_policy_select_batch_tokens
        secs = fn(self.h, _ptr(tk), _ptr(off), n, _ptr(idx), _ptr(br), _ptr(ma))
        return idx, br, ma, secs


# ---- prefix_hash policy + consistent hash ring (oracle/prefix_hash.h) ----
PREFIX_BRANCHES = ["no_healthy_workers", "no_tokens", "ring_hit", "load_balance_walk", "fallback_least_load"]   # prefix_hash.rs:70-83


class HashRing:                              # model_gateway/src/worker/hash_ring.rs
    def __init__(self, urls):
        self.urls = list(urls)
        enc = [u.encode() for u in self.urls]
        arr = (C.c_char_p * max(len(enc), 1))(*enc)
        self._h = lib().orc_ring_new(arr, len(enc))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ring_free(self._h)
            self._h = None

    @staticmethod
    def hash_position(s: str) -> int:        # :78-86
        b = s.encode("utf-8")
        return lib().orc_ring_hash_position(b, len(b))

    def is_empty(self):
        return len(self) == 0

    def __len__(self):
        return lib().orc_ring_len(self._h)

    def worker_count(self):
        return lib().orc_ring_worker_count(self._h)

    def entries(self):
        n = len(self)
        pos = np.zeros(max(n, 1), np.uint64)
        url = np.zeros(max(n, 1), np.uint32)
        lib().orc_ring_entries(self._h, _ptr(pos), _ptr(url), n)
        return pos[:n], url[:n]

    def find_healthy_url(self, key: str, is_healthy):   # :102-134
        if not self.urls:
            return None
        h = np.array([1 if is_healthy(u) else 0 for u in self.urls], np.uint8)
        b = key.encode("utf-8")
        i = lib().orc_ring_find_healthy(self._h, b, len(b), _ptr(h))
        return None if i < 0 else self.urls[i]


class PrefixHashPolicy:                      # model_gateway/src/policies/prefix_hash.rs
    def __init__(self, prefix_token_count=256, load_factor=1.25):
        self.prefix_token_count, self.load_factor = int(prefix_token_count), float(load_factor)

    def name(self):
        return "prefix_hash"

    def compute_prefix_hash(self, tokens) -> int:
        t = _u32(tokens)
        return lib().orc_prefix_hash(_ptr(t), t.size, self.prefix_token_count)

    def load_ok(self, worker_load, total_load, num_workers) -> bool:
        return bool(lib().orc_prefix_load_ok(self.load_factor, int(worker_load), int(total_load), int(num_workers)))

    def select_batch(self, urls, loads, healthy, ring, tokens, offsets, has_tokens=True):
        """One batch against one fleet snapshot → (idx, branch, seconds).  ring None = info.hash_ring None."""
        enc = [u.encode() for u in urls]
        arr = (C.c_char_p * max(len(enc), 1))(*enc)
        ld, hl = _u64(loads), np.ascontiguousarray(np.asarray(healthy, np.uint8))
        tk, off = _u32(tokens), _u64(offsets)
        n = off.size - 1
        idx = np.full(max(n, 1), -1, np.int32)
        br = np.zeros(max(n, 1), np.uint8)
        secs = lib().orc_prefix_select_batch(self.prefix_token_count, self.load_factor, arr, _ptr(ld), _ptr(hl), len(enc),
                                             ring._h if ring is not None else None, _ptr(tk), _ptr(off), n, 1 if has_tokens else 0, _ptr(idx), _ptr(br))
        return idx[:n], br[:n], secs

    def select_worker(self, urls, loads, healthy, ring, tokens):
        """select_worker_impl (:203-222): tokens None = info.tokens None.  → (idx or None, branch name)"""
        t = _u32(tokens if tokens is not None else [])
        idx, br, _ = self.select_batch(urls, loads, healthy, ring, t, np.array([0, t.size], np.uint64), has_tokens=tokens is not None)
        return (None if idx[0] < 0 else int(idx[0])), PREFIX_BRANCHES[int(br[0])]


class PowerOfTwoPolicy:                     # model_gateway/src/policies/power_of_two.rs
    """select over an explicit draw stream (oracle/power_of_two.h): request i of a batch uses draws 2i and 2i + 1 of `seed`."""

    def __init__(self):
        self.cached = {}                     # url → effective_token_usage (update_loads, :129-135)

    def name(self):
        return "power_of_two"

    @staticmethod
    def effective_token_usage(token_usage_per_dp_rank) -> float:   # protocols worker.rs:1039-1044
        a = np.ascontiguousarray(np.asarray(token_usage_per_dp_rank, np.float64))
        return float(lib().orc_p2c_effective_token_usage(_ptr(a) if a.size else None, a.size))

    def update_loads(self, loads: dict):
        self.cached.update({k: float(v) for k, v in loads.items()})

    def select_batch(self, urls, loads, healthy, circuit, n, seed):
        """→ (idx int32[n], pairs int32[n, 2], metric uint8[n], processed uint64[W]) against ONE fleet snapshot."""
        enc = [u.encode() for u in urls]
        arr = (C.c_char_p * max(len(enc), 1))(*enc)
        cu = [k.encode() for k in self.cached]
        carr = (C.c_char_p * max(len(cu), 1))(*cu)
        cv = np.ascontiguousarray(np.asarray(list(self.cached.values()) or [0.0], np.float64))
        ld, hl = _u64(loads), np.ascontiguousarray(np.asarray(healthy, np.uint8))
        ck = np.ascontiguousarray(np.asarray(circuit if circuit is not None else [1] * len(enc), np.uint8))
        idx = np.full(max(n, 1), -1, np.int32)
        pairs = np.full((max(n, 1), 2), -1, np.int32)
        metric = np.zeros(max(n, 1), np.uint8)
        proc = np.zeros(max(len(enc), 1), np.uint64)
        lib().orc_p2c_select_batch(arr, _ptr(ld), _ptr(hl), _ptr(ck), len(enc), carr, _ptr(cv), len(cu), int(seed) & (2**64 - 1), n, _ptr(idx), _ptr(pairs),
                                   _ptr(metric), _ptr(proc))
        return idx[:n], pairs[:n], metric[:n], proc[:len(enc)]

    def select_worker(self, urls, loads, healthy, circuit=None, seed=0):
        idx, pairs, metric, _ = self.select_batch(urls, loads, healthy, circuit, 1, seed)
        return None if idx[0] < 0 else int(idx[0])
