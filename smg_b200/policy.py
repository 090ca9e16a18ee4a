"""Host-side mirror of the reference's plugin surface for the cache-aware path, over the smgx C ABI.

Names, argument meaning and error behaviour follow the reference so the parity tests read like its own tests:
  * CacheAwareConfig / SelectWorkerInfo / LoadBalancingPolicy   model_gateway/src/policies/mod.rs:43-175
  * CacheAwarePolicy                                            model_gateway/src/policies/cache_aware.rs
  * PolicyFactory                                               model_gateway/src/policies/factory.rs:17-93
  * PositionalIndexer                                           crates/kv_index/src/event_tree.rs:257-760
  * KvEventMonitor (get_indexer / block_size / set_block_size)  model_gateway/src/worker/kv_event_monitor.rs
  * the Worker scalars the path reads                           model_gateway/src/worker/worker.rs:109-230
Every decision comes out of the CUDA kernels; this file only marshals arrays.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

UNKNOWN_MODEL_ID = "unknown"


def normalize_model_key(model_id: str) -> str:  # policies/mod.rs:151-157
    return model_id if model_id else UNKNOWN_MODEL_ID


@dataclass
class CacheAwareConfig:  # policies/mod.rs:94-117
    cache_threshold: float = 0.5
    balance_abs_threshold: int = 32
    balance_rel_threshold: float = 1.1
    eviction_interval_secs: int = 30
    max_tree_size: int = 10000
    block_size: int = 16


@dataclass
class SelectWorkerInfo:  # policies/mod.rs:161-175 (headers are not read on this path; hash_ring only by prefix_hash)
    request_text: Optional[str] = None
    tokens: Optional[Sequence[int]] = None
    hash_ring: Optional["HashRing"] = None


class BasicWorker:
    """The scalars CacheAwarePolicy reads through `trait Worker` (worker/worker.rs:114,151-153,187,208,217)."""

    def __init__(self, url: str, model_id: str = ""):
        self._url, self._model_id = url, model_id
        self._load, self._healthy, self._circuit_ok, self._processed = 0, True, True, 0

    def url(self): return self._url
    def model_id(self): return self._model_id
    def load(self): return self._load
    def is_healthy(self): return self._healthy
    def circuit_breaker_can_execute(self): return self._circuit_ok
    def increment_processed(self): self._processed = getattr(self, "_processed", 0) + 1
    def processed(self): return getattr(self, "_processed", 0)
    def processed(self): return self._processed
    def increment_load(self): self._load += 1
    def decrement_load(self): self._load = max(0, self._load - 1)
    def set_load(self, v): self._load = int(v)
    def set_healthy(self, ok: bool): self._healthy = bool(ok)      # set_status(Ready) ⇔ True; any other status ⇔ False
    def set_circuit_ok(self, ok: bool): self._circuit_ok = bool(ok)


def _u32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class _Handle:
    """Owns one smgx_policy*."""

    def __init__(self, cfg: CacheAwareConfig, device_id: int, max_batch: int = 0, max_tokens_per_request: int = 0, tree_batch_mode: int = 0):
        self.L = _lib.load()
        c = _lib.Config()
        self.L.smgx_default_config(C.byref(c))
        c.cache_threshold, c.balance_abs_threshold = cfg.cache_threshold, cfg.balance_abs_threshold
        c.balance_rel_threshold, c.eviction_interval_secs = cfg.balance_rel_threshold, cfg.eviction_interval_secs
        c.max_tree_size, c.block_size = cfg.max_tree_size, cfg.block_size
        c.device_id, c.max_batch, c.max_tokens_per_request = device_id, max_batch, max_tokens_per_request
        c.tree_batch_mode = tree_batch_mode
        err = _lib.new_err()
        self.p = self.L.smgx_policy_create(C.byref(c), C.byref(err))
        if not self.p:
            _lib.check(_lib.UNKNOWN_ERROR if not err.value else _lib.DEVICE_ERROR, err)
        self.device_id = device_id

    def close(self):
        if getattr(self, "p", None):
            self.L.smgx_policy_free(self.p)
            self.p = None

    def __del__(self):
        self.close()

    def call(self, fn, *args):
        err = _lib.new_err()
        code = getattr(self.L, fn)(self.p, *args, C.byref(err))
        _lib.check(code, err)


class ApplyError(Exception):  # kv_index::ApplyError (event_tree.rs:88-106)
    pass


class PositionalIndexer:
    """kv_index::PositionalIndexer bound to (policy handle, model).  Blocks are (seq_hash, content_hash) pairs =
    StoredBlock (event_tree.rs:81-86); the caller-owned WorkerBlockMap lives inside the library per worker id."""

    def __init__(self, handle: _Handle, model: str, jump_size: int = 32):
        if jump_size <= 0:
            raise ValueError("jump_size must be greater than 0")  # event_tree.rs:280
        self.h, self.model = handle, model.encode()
        self.h.call("smgx_indexer_create", self.model, jump_size)

    @classmethod
    def standalone(cls, jump_size: int = 32, device_id: int = 0):
        """An indexer with a private policy handle (the reference constructs PositionalIndexer::new(jump) directly)."""
        if jump_size <= 0:
            raise ValueError("jump_size must be greater than 0")
        return cls(_Handle(CacheAwareConfig(eviction_interval_secs=0), device_id), UNKNOWN_MODEL_ID, jump_size)

    def intern_worker(self, url: str) -> int:
        out = C.c_uint32()
        self.h.call("smgx_indexer_intern_worker", self.model, url.encode(), C.byref(out))
        return out.value

    def worker_id(self, url: str) -> Optional[int]:
        out = C.c_int64()
        self.h.call("smgx_indexer_worker_id", self.model, url.encode(), C.byref(out))
        return None if out.value < 0 else out.value

    def apply_stored(self, worker_id: int, blocks, parent: Optional[int] = None):
        seq, con = _u64([b[0] for b in blocks]), _u64([b[1] for b in blocks])
        par = C.c_uint64(parent) if parent is not None else None
        try:
            self.h.call("smgx_indexer_apply_stored", self.model, worker_id, _p(seq), _p(con), len(blocks),
                        C.cast(C.byref(par), C.c_void_p) if par is not None else None)
        except _lib.SmgxError as e:
            if e.code == _lib.WORKER_NOT_TRACKED:
                raise ApplyError("WorkerNotTracked: " + e.msg)
            if e.code == _lib.PARENT_BLOCK_NOT_FOUND:
                raise ApplyError("ParentBlockNotFound: " + e.msg)
            raise

    def apply_stored_tokens(self, worker_id: int, seq_hashes, token_ids, block_size: int, parent: Optional[int] = None):
        seq, tok = _u64(seq_hashes), _u32(token_ids)
        par = C.c_uint64(parent) if parent is not None else None
        try:
            self.h.call("smgx_indexer_apply_stored_tokens", self.model, worker_id, _p(seq), _p(tok), block_size, seq.size,
                        C.cast(C.byref(par), C.c_void_p) if par is not None else None)
        except _lib.SmgxError as e:
            if e.code == _lib.WORKER_NOT_TRACKED:
                raise ApplyError("WorkerNotTracked: " + e.msg)
            if e.code == _lib.PARENT_BLOCK_NOT_FOUND:
                raise ApplyError("ParentBlockNotFound: " + e.msg)
            raise

    def apply_removed(self, worker_id: int, seq_hashes):
        s = _u64(seq_hashes)
        self.h.call("smgx_indexer_apply_removed", self.model, worker_id, _p(s), s.size)

    def apply_cleared(self, worker_id: int):
        self.h.call("smgx_indexer_apply_cleared", self.model, worker_id)

    def remove_worker(self, worker_id: int):
        self.h.call("smgx_indexer_remove_worker", self.model, worker_id)

    def current_size(self) -> int:
        out = C.c_uint64()
        self.h.call("smgx_indexer_current_size", self.model, C.byref(out))
        return out.value

    def entry_count(self) -> int:
        out = C.c_uint64()
        self.h.call("smgx_indexer_entry_count", self.model, C.byref(out))
        return out.value

    def find_matches(self, content_hashes, early_exit: bool = False):
        """→ (scores, tree_sizes) dicts keyed by worker id, like OverlapScores (event_tree.rs:112-118).  GPU kernel."""
        hs = _u64(content_hashes)
        cap = 2048
        sc, ts, nw = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), C.c_uint32()
        self.h.call("smgx_indexer_find_matches", self.model, _p(hs), hs.size, 1 if early_exit else 0, _p(sc), _p(ts), cap, C.byref(nw))
        ids = [i for i in range(nw.value) if sc[i] > 0]
        return {i: int(sc[i]) for i in ids}, {i: int(ts[i]) for i in ids}


class KvEventMonitor:
    """The slice of worker::KvEventMonitor the policy reads: per-model indexers and learned block sizes."""

    def __init__(self, policy: "CacheAwarePolicy", default_block_size: Optional[int] = None):
        self.policy, self.default_block_size, self.indexers = policy, default_block_size, {}
        self._learned = {}

    def create_indexer(self, model: str, jump_size: int = 64) -> PositionalIndexer:  # DEFAULT_JUMP_SIZE kv_event_monitor.rs:31
        ix = PositionalIndexer(self.policy._h, model, jump_size)
        self.indexers[model] = ix
        return ix

    def get_indexer(self, model: str) -> Optional[PositionalIndexer]:
        return self.indexers.get(model)

    def set_block_size(self, model: str, block_size: int):
        self.policy._h.call("smgx_indexer_set_block_size", model.encode(), block_size)

    def apply_events(self, model: str, worker_id: int, events):
        """One KvEventBatch of a worker's stream (kv_event_monitor.rs:513-517, apply_event :525-597) through smgx_kv_events_apply.
        `events`: dicts shaped like the proto — {"stored": {"blocks": [{"block_hash", "token_ids", "block_size"}...],
        "parent_block_hash": int | None}}, {"removed": {"block_hashes": [...]}} or {"cleared": {}}.  The first stored block with
        block_size > 0 teaches the model's block size once (learn_block_size :270-296).  Returns the number of fresh-chain fallbacks."""
        evs, hashes, toks, offs = [], [], [], [0]
        for ev in events:
            if "stored" in ev:
                st = ev["stored"]
                blocks = st.get("blocks", [])
                if blocks and not self._learned.get(model) and blocks[0].get("block_size", 0) > 0:
                    self.set_block_size(model, int(blocks[0]["block_size"]))
                    self._learned[model] = True
                parent = st.get("parent_block_hash")
                evs.append((0, len(hashes), len(blocks), 0 if parent is None else int(parent), 0 if parent is None else 1))
                for b in blocks:
                    hashes.append(int(b["block_hash"]))
                    toks.extend(int(t) for t in b.get("token_ids", []))
                    offs.append(len(toks))
            elif "removed" in ev:
                hs = [int(h) for h in ev["removed"].get("block_hashes", [])]
                evs.append((1, len(hashes), len(hs), 0, 0))
                for h in hs:
                    hashes.append(h)
                    offs.append(len(toks))
            elif "cleared" in ev:
                evs.append((2, 0, 0, 0, 0))
        arr = (_lib.KvEvent * max(len(evs), 1))()
        for i, (k, fb, nb, par, hp) in enumerate(evs):
            arr[i].kind, arr[i].worker_id, arr[i].first_block, arr[i].n_blocks = k, worker_id, fb, nb
            arr[i].parent_block_hash, arr[i].has_parent = par, hp
        h64 = np.ascontiguousarray(np.asarray([x if x < 2**63 else x - 2**64 for x in hashes] or [0], dtype=np.int64))
        t32, o32 = _u32(toks), _u32(offs)
        fallbacks = C.c_uint32()
        self.policy._h.call("smgx_kv_events_apply", model.encode(), C.cast(arr, C.c_void_p), len(evs), _p(h64), _p(o32), _p(t32) if t32.size else None,
                            len(hashes), C.byref(fallbacks))
        return fallbacks.value


class TokenMatchResult:  # kv_index::PrefixMatchResult (token_tree.rs:137-144)
    def __init__(self, tenant, matched, inp):
        self.tenant, self.matched_token_count, self.input_token_count = tenant, matched, inp


LRU, LFU, FIFO, MRU, FILO, PRIORITY = range(6)   # kv_index::EvictionPolicy (token_tree.rs:60-75)


class TokenTree:
    """kv_index::TokenTree bound to (policy handle, model): mutations on the host-authoritative tree inside the library,
    match_prefix_with_counts on the GPU mirror."""

    def __init__(self, handle: _Handle, model: str = UNKNOWN_MODEL_ID, policy: int = LRU):
        self.h, self.model = handle, model.encode()
        self.h.call("smgx_tree_create", self.model, policy)

    @classmethod
    def standalone(cls, policy: int = LRU, device_id: int = 0):
        return cls(_Handle(CacheAwareConfig(eviction_interval_secs=0), device_id), UNKNOWN_MODEL_ID, policy)

    def insert_tokens(self, tokens, tenant: str):
        t = _u32(tokens)
        self.h.call("smgx_tree_insert_tokens", self.model, _p(t), t.size, tenant.encode())

    def match_prefix_with_counts(self, tokens) -> TokenMatchResult:
        t = _u32(tokens)
        m, n = C.c_uint32(), C.c_uint32()
        buf = C.create_string_buffer(1024)
        self.h.call("smgx_tree_match_tokens", self.model, _p(t), t.size, C.byref(m), C.byref(n), C.cast(buf, C.c_void_p), 1024)
        return TokenMatchResult(buf.value.decode(), m.value, n.value)

    def evict_tenant(self, tenant: str, max_tokens: int):
        self.h.call("smgx_tree_evict_tenant", self.model, tenant.encode(), max_tokens)

    def evict_tenant_by_size(self, max_size: int):
        self.h.call("smgx_evict_cache", max_size)

    def tenant_token_size(self, tenant: str) -> int:
        out = C.c_uint64()
        self.h.call("smgx_tree_tenant_size", self.model, tenant.encode(), C.byref(out))
        return out.value

    def clear(self):
        self.h.call("smgx_tree_clear", self.model)

    def entries(self):
        out = C.c_void_p()
        self.h.call("smgx_tree_entries", self.model, C.byref(out))
        text = C.cast(out, C.c_char_p).value.decode()
        self.h.L.smgx_free_string(out)
        res = []
        for line in text.split("\n"):
            if not line:
                continue
            toks, tens = line.split("|")
            res.append(([int(x) for x in toks.split(",")] if toks else [],
                        [(kv.rsplit("=", 1)[0], int(kv.rsplit("=", 1)[1])) for kv in tens.split(";") if kv]))
        return res


class StringMatchResult:  # kv_index::PrefixMatchResult of the string tree (string_tree.rs:47-54)
    def __init__(self, tenant, matched, inp):
        self.tenant, self.matched_char_count, self.input_char_count = tenant, matched, inp


class Tree:
    """kv_index::Tree (char-level string tree, HTTP text routing) bound to (policy handle, model): mutations on the
    host-authoritative tree inside the library, match_prefix_with_counts on the GPU mirror."""

    def __init__(self, handle: _Handle, model: str = UNKNOWN_MODEL_ID):
        self.h, self.model = handle, model.encode()

    @classmethod
    def standalone(cls, device_id: int = 0):
        return cls(_Handle(CacheAwareConfig(eviction_interval_secs=0), device_id), UNKNOWN_MODEL_ID)

    @staticmethod
    def _b(text: str):
        b = text.encode("utf-8")
        return (np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(0, np.uint8)), len(b)

    def insert_text(self, text: str, tenant: str):
        a, n = self._b(text)
        self.h.call("smgx_stree_insert_text", self.model, _p(a), n, tenant.encode())

    def match_prefix_with_counts(self, text: str) -> StringMatchResult:
        a, n = self._b(text)
        m, k = C.c_uint32(), C.c_uint32()
        buf = C.create_string_buffer(1024)
        self.h.call("smgx_stree_match", self.model, _p(a), n, C.byref(m), C.byref(k), C.cast(buf, C.c_void_p), 1024)
        return StringMatchResult(buf.value.decode(), m.value, k.value)

    def prefix_match_tenant(self, text: str, tenant: str) -> str:
        a, n = self._b(text)
        out = C.c_uint32()
        self.h.call("smgx_stree_prefix_match_tenant", self.model, _p(a), n, tenant.encode(), C.byref(out))
        return text.encode("utf-8")[: out.value].decode("utf-8")

    def evict_tenant_by_size(self, max_size: int):
        self.h.call("smgx_evict_cache", max_size)

    def _sizes(self, maintained: int):
        out = C.c_void_p()
        self.h.call("smgx_stree_sizes", self.model, maintained, C.byref(out))
        text = C.cast(out, C.c_char_p).value.decode()
        self.h.L.smgx_free_string(out)
        res = {}
        for line in text.split("\n"):
            if line:
                k, v = line.rsplit("=", 1)
                res[k] = int(v)
        return res

    def get_used_size_per_tenant(self):
        return self._sizes(0)

    def get_tenant_char_count(self):
        return self._sizes(1)

    def node_count(self) -> int:
        out = C.c_uint64()
        self.h.call("smgx_stree_node_count", self.model, C.byref(out))
        return out.value

    def clear(self):
        self.h.call("smgx_stree_clear", self.model)

    # -- mesh wire format: kv_index::snapshot::TreeSnapshot, bincode (snapshot.rs; string_tree.rs:1052-1578) --
    def snapshot_bytes(self) -> bytes:  # Tree::snapshot().to_bytes()
        out, ln = C.c_void_p(), C.c_uint64()
        self.h.call("smgx_stree_snapshot", self.model, C.byref(out), C.byref(ln))
        raw = C.string_at(out, ln.value)
        self.h.L.smgx_free_string(out)
        return raw

    @classmethod
    def from_snapshot_bytes(cls, data: bytes, device_id: int = 0):  # TreeSnapshot::from_bytes + Tree::from_snapshot (:1228)
        t = cls.standalone(device_id)
        t.h.call("smgx_stree_load_snapshot", t.model, data, len(data))
        return t

    def merge_snapshot_bytes(self, data: bytes):  # Tree::merge_snapshot (:1318)
        self.h.call("smgx_stree_merge_snapshot", self.model, data, len(data))

    def entries(self):
        out, ln = C.c_void_p(), C.c_uint64()
        self.h.call("smgx_stree_entries", self.model, C.byref(out), C.byref(ln))
        raw = C.string_at(out, ln.value).decode("utf-8")
        self.h.L.smgx_free_string(out)
        res = []
        for rec in raw.split("\x1e"):
            if not rec:
                continue
            path, tens = rec.split("\x1f")
            res.append((path, [(kv.rsplit("=", 1)[0], int(kv.rsplit("=", 1)[1])) for kv in tens.split(";") if kv]))
        return res


class TiktokenTokenizer:
    """tokenizer::TiktokenTokenizer (crates/tokenizer/src/tiktoken.rs:132) on the GPU, bound to (policy handle, model).
    encode() keeps the reference's behaviour of always recognising special-token strings (tiktoken.rs:444-462)."""

    def __init__(self, handle: _Handle, model: str, path: str, special_tokens=None):
        self.h, self.model = handle, model.encode()
        sp = list((special_tokens or {}).items())
        strs = (C.c_char_p * max(len(sp), 1))(*[k.encode() for k, _ in sp]) if sp else None
        ids = _u32([v for _, v in sp])
        self.h.call("smgx_tokenizer_load_tiktoken_file", self.model, path.encode(), strs, _p(ids), len(sp))

    @staticmethod
    def _ragged(texts):
        blobs = [t.encode("utf-8") for t in texts]
        offsets = np.zeros(len(blobs) + 1, dtype=np.uint32)
        np.cumsum([len(b) for b in blobs], out=offsets[1:])
        data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8).copy()
        return data, offsets

    def encode_batch(self, texts, add_special_tokens: bool = False):
        data, offsets = self._ragged(texts)
        n = len(texts)
        cap = int(offsets[-1]) + 1
        out, toff = np.zeros(cap, np.uint32), np.zeros(n + 1, np.uint32)
        self.h.call("smgx_tokenize_batch", self.model, _p(data), _p(offsets), n, _p(out), _p(toff), cap)
        return [out[toff[i]:toff[i + 1]].tolist() for i in range(n)]

    def encode(self, text: str, add_special_tokens: bool = False):
        return self.encode_batch([text])[0]


LLAMA3_SPLIT_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")


def _byte_level_decoder():
    """Inverse of the GPT-2 bytes_to_unicode table used by `tokenizers`' ByteLevel pre-tokenizer (pre_tokenizers/byte_level.rs)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {chr(c): b for b, c in zip(bs, cs)}


class HuggingFaceTokenizer(TiktokenTokenizer):
    """tokenizer::HuggingFaceTokenizer (crates/tokenizer/src/huggingface.rs) for the byte-level BPE family the GPU path covers:
    tokenizer.json with no normalizer, Split(Llama-3 / cl100k regex, Isolated) + ByteLevel(use_regex=False), model BPE without
    dropout / unk / byte_fallback.  Anything else is refused loudly — there is no CPU tokenizer behind this."""

    def __init__(self, handle: _Handle, model: str, path: str):   # noqa: D401 — does not call the tiktoken loader
        import json
        self.h, self.model = handle, model.encode()
        j = json.load(open(path, encoding="utf-8"))
        mdl = j.get("model") or {}
        if mdl.get("type") != "BPE" or mdl.get("dropout") or mdl.get("unk_token") or mdl.get("byte_fallback") or \
                mdl.get("continuing_subword_prefix") or mdl.get("end_of_word_suffix"):
            raise ValueError("tokenizer.json: only plain byte-level BPE models are supported on the GPU path")
        if j.get("normalizer") is not None:
            raise ValueError("tokenizer.json: normalizers are not supported on the GPU path")
        pre = j.get("pre_tokenizer") or {}
        steps = pre.get("pretokenizers", [pre]) if pre.get("type") == "Sequence" else [pre]
        kinds = [s.get("type") for s in steps]
        split = [s for s in steps if s.get("type") == "Split"]
        bl = [s for s in steps if s.get("type") == "ByteLevel"]
        if kinds != ["Split", "ByteLevel"] or split[0].get("pattern", {}).get("Regex") != LLAMA3_SPLIT_PATTERN or \
                str(split[0].get("behavior")).lower() != "isolated" or split[0].get("invert") or bl[0].get("use_regex") or bl[0].get("add_prefix_space"):
            raise ValueError("tokenizer.json: pre_tokenizer must be Split(Llama-3 regex, Isolated) + ByteLevel(use_regex=False)")
        dec = _byte_level_decoder()
        vocab = mdl["vocab"]
        toks = [bytes(dec[c] for c in s) for s in vocab]
        ids = _u32(list(vocab.values()))
        merges = []
        for mg in mdl.get("merges", []):
            a, b = mg.split(" ") if isinstance(mg, str) else mg
            merges += [vocab[a], vocab[b]]
        offs = np.zeros(len(toks) + 1, dtype=np.uint32)
        np.cumsum([len(t) for t in toks], out=offs[1:])
        blob = np.frombuffer(b"".join(toks), dtype=np.uint8).copy()
        added = [(a["content"], a["id"]) for a in j.get("added_tokens", [])]
        for a in j.get("added_tokens", []):
            if a.get("lstrip") or a.get("rstrip") or a.get("single_word") or a.get("normalized"):
                raise ValueError("tokenizer.json: added tokens with lstrip / rstrip / single_word / normalized are not supported")
        strs = (C.c_char_p * max(len(added), 1))(*[k.encode() for k, _ in added]) if added else None
        sp_ids = _u32([v for _, v in added])
        mg = _u32(merges)
        self.h.call("smgx_tokenizer_load_bpe_merges", self.model, _p(blob), _p(offs), _p(ids), len(toks), _p(mg), len(merges) // 2,
                    1 if mdl.get("ignore_merges") else 0, strs, _p(sp_ids), len(added))


TREE_BATCH_MODES = {"sequential": 0, "snapshot": 1}   # smgx_tree_batch_mode (include/smgx.h)


class CacheAwarePolicy:
    """policies::CacheAwarePolicy on the GPU.  select_worker keeps the reference signature; select_worker_batch is the
    batched form the host batcher uses (all requests see one fleet snapshot)."""

    def __init__(self, config: Optional[CacheAwareConfig] = None, device_id: int = 0, max_batch: int = 0, max_tokens_per_request: int = 0,
                 tree_batch_mode: str = "sequential"):
        self.config = config or CacheAwareConfig()
        self._h = _Handle(self.config, device_id, max_batch, max_tokens_per_request, TREE_BATCH_MODES[tree_batch_mode])
        self._slices = {}   # model → tuple(urls) currently registered
        self._monitor = None

    @classmethod
    def with_config(cls, config: CacheAwareConfig, **kw):
        return cls(config, **kw)

    def close(self):
        self._h.close()

    # -- LoadBalancingPolicy ------------------------------------------------------------------------------------
    def name(self) -> str:
        return _lib.load().smgx_policy_name().decode()

    def needs_request_text(self) -> bool:  # cache_aware.rs:708-710
        return True

    def on_request_complete(self, worker_url: str, success: bool):  # cache_aware.rs:692-702 (no state)
        return None

    # -- fleet ---------------------------------------------------------------------------------------------------
    def _model_of(self, workers: Sequence[BasicWorker]) -> str:
        for w in workers:  # model of the first healthy worker (cache_aware.rs:659); any worker when none is healthy
            if w.is_healthy() and w.circuit_breaker_can_execute():
                return normalize_model_key(w.model_id())
        return normalize_model_key(workers[0].model_id()) if workers else UNKNOWN_MODEL_ID

    def _register(self, model: str, urls):
        arr = (C.c_char_p * len(urls))(*[u.encode() for u in urls])
        self._h.call("smgx_set_workers", model.encode(), arr, len(urls))
        self._slices[model] = tuple(urls)

    def init_workers(self, workers: Sequence[BasicWorker]):  # cache_aware.rs:219-249
        by_model = {}
        for w in workers:
            by_model.setdefault(normalize_model_key(w.model_id()), []).append(w.url())
        for model, urls in by_model.items():
            self._register(model, urls)

    def add_worker(self, worker: BasicWorker):  # cache_aware.rs:252-266
        self._h.call("smgx_add_worker", normalize_model_key(worker.model_id()).encode(), worker.url().encode())

    def remove_worker_by_url(self, url: str):  # cache_aware.rs:302-308 (no-op)
        self._h.call("smgx_remove_worker", b"", url.encode())

    def _push_fleet(self, workers: Sequence[BasicWorker]) -> bytes:
        model = self._model_of(workers)
        urls = tuple(w.url() for w in workers)
        if self._slices.get(model) != urls:
            self._register(model, list(urls))
        loads = _u64([w.load() for w in workers])
        healthy = np.ascontiguousarray(np.asarray([1 if w.is_healthy() else 0 for w in workers], dtype=np.uint8))
        circuit = np.ascontiguousarray(np.asarray([1 if w.circuit_breaker_can_execute() else 0 for w in workers], dtype=np.uint8))
        self._h.call("smgx_set_fleet_state", model.encode(), _p(loads), _p(healthy), _p(circuit), len(workers))
        return model.encode()

    # -- KV events -----------------------------------------------------------------------------------------------
    def kv_event_monitor(self, default_block_size: Optional[int] = None) -> KvEventMonitor:
        return KvEventMonitor(self, default_block_size)

    def set_kv_event_monitor(self, monitor: Optional[KvEventMonitor]):  # cache_aware.rs:213-215
        self._monitor = monitor
        self._h.call("smgx_set_kv_event_monitor", 1 if monitor is not None else 0)

    # -- the pick ------------------------------------------------------------------------------------------------
    def select_worker(self, workers: Sequence[BasicWorker], info: SelectWorkerInfo) -> Optional[int]:
        """LoadBalancingPolicy::select_worker (cache_aware.rs:648): index into `workers` or None."""
        if not workers:
            return None
        if info.tokens is None:   # HTTP routers: string tree on request_text.unwrap_or("") (cache_aware.rs:688-689)
            idx, _ = self.select_worker_batch_request_text(workers, [info.request_text or ""])
        else:
            idx, _ = self.select_worker_batch(workers, [info.tokens])
        i = int(idx[0])
        if i >= 0:
            workers[i]._processed += 1  # mirror of increment_processed(); the library keeps the authoritative counters
        return None if i < 0 else i

    def select_worker_batch(self, workers: Sequence[BasicWorker], requests=None, tokens=None, offsets=None, want_info: bool = True):
        """Batch of token requests against one fleet snapshot.  Either `requests` (list of token lists) or a ragged
        (tokens u32, offsets u32[n+1]) pair.  → (worker_idx int32[n] with -1 = None, info structured array or None)."""
        model = self._push_fleet(workers)
        if requests is not None:
            lens = [len(r) for r in requests]
            offsets = np.zeros(len(requests) + 1, dtype=np.uint32)
            np.cumsum(lens, out=offsets[1:])
            tokens = _u32(np.concatenate([np.asarray(r, dtype=np.uint32) for r in requests]) if sum(lens) else [])
        tokens, offsets = _u32(tokens), _u32(offsets)
        n = offsets.size - 1
        out = np.full(max(n, 1), -1, dtype=np.int32)
        info = (_lib.DecisionInfo * max(n, 1))() if want_info else None
        tok_ptr = _p(tokens) if tokens.size else C.cast(C.create_string_buffer(4), C.c_void_p)
        self._h.call("smgx_select_batch_tokens", model, tok_ptr, _p(offsets), n, _p(out), C.cast(info, C.c_void_p) if info is not None else None)
        return out[:n], (info if info is None else [info[i] for i in range(n)])

    def select_worker_batch_request_text(self, workers: Sequence[BasicWorker], texts=None, want_info: bool = True, data=None, offsets=None):
        """Batch of HTTP-mode requests (request_text, no tokens) against one fleet snapshot: string-tree walk + pick on the
        GPU (select_worker_with_text, cache_aware.rs:907-974).  Either `texts` (list of str) or a ragged UTF-8 pair
        (data uint8, offsets uint32[n+1]).  → (worker_idx int32[n], info list with char counts)."""
        model = self._push_fleet(workers)
        if texts is not None:
            data, offsets = TiktokenTokenizer._ragged(texts)
        n = offsets.size - 1
        out = np.full(max(n, 1), -1, dtype=np.int32)
        info = (_lib.DecisionInfo * max(n, 1))() if want_info else None
        self._h.call("smgx_select_batch_request_text", model, _p(data), _p(offsets), n, _p(out), C.cast(info, C.c_void_p) if info is not None else None)
        return out[:n], (info if info is None else [info[i] for i in range(n)])

    # -- mesh path hashes / hash_index (crates/mesh/src/hash.rs; cache_aware.rs:95-101) ---------------------------------
    def hash_token_paths(self, requests):
        lens = [len(r) for r in requests]
        offsets = np.zeros(len(requests) + 1, dtype=np.uint32)
        np.cumsum(lens, out=offsets[1:])
        tokens = _u32(np.concatenate([np.asarray(r, dtype=np.uint32) for r in requests]) if sum(lens) else [])
        out = np.zeros(max(len(requests), 1), np.uint64)
        self._h.call("smgx_hash_token_paths", _p(tokens), _p(offsets), len(requests), _p(out))
        return out[:len(requests)]

    def hash_node_paths(self, texts):
        data, offsets = TiktokenTokenizer._ragged(texts)
        out = np.zeros(max(len(texts), 1), np.uint64)
        self._h.call("smgx_hash_node_paths", _p(data), _p(offsets), len(texts), _p(out))
        return out[:len(texts)]

    def hash_index_size(self, kind: str = "tokens", model: str = UNKNOWN_MODEL_ID) -> int:
        out = C.c_uint64()
        self._h.call("smgx_hash_index_size", model.encode(), 1 if kind == "text" else 0, C.byref(out))
        return out.value

    def hash_index_get(self, path_hash: int, kind: str = "tokens", model: str = UNKNOWN_MODEL_ID):
        """hash_index[model].token_tree / .string_tree lookup → matched-prefix copy (list of ids / str) or None."""
        nb, found = C.c_uint32(), C.c_int()
        tk = 1 if kind == "text" else 0
        self._h.call("smgx_hash_index_get", model.encode(), tk, int(path_hash), None, 0, C.byref(nb), C.byref(found))
        if not found.value:
            return None
        buf = np.zeros(max(nb.value, 1), np.uint8)
        self._h.call("smgx_hash_index_get", model.encode(), tk, int(path_hash), _p(buf), nb.value, C.byref(nb), C.byref(found))
        raw = buf[:nb.value].tobytes()
        return raw.decode("utf-8") if tk else np.frombuffer(raw, dtype=np.uint32).tolist()

    # -- TreeHandle (cache_aware.rs:454-645) ----------------------------------------------------------------------------
    def apply_known_remote_insert(self, model_id: str, tree_kind: str, node_hash: int, worker_url: str) -> bool:
        known = C.c_int()
        self._h.call("smgx_tree_apply_known_remote_insert", (model_id or "").encode(), 1 if tree_kind == "token" else 0, int(node_hash), worker_url.encode(), C.byref(known))
        return bool(known.value)

    def apply_repair_page(self, model_id: str, tree_kind: str, entries) -> int:
        """entries: [("string", path: str, [(tenant, epoch)]) | ("token", tokens, [(tenant, epoch)])] → entries applied."""
        arr = (_lib.RepairEntry * max(len(entries), 1))()
        keep = []
        for i, (kind, path, tenants) in enumerate(entries):
            data = np.frombuffer(path.encode(), dtype=np.uint8).copy() if kind == "string" else _u32(path)
            names = (C.c_char_p * max(len(tenants), 1))(*[t.encode() for t, _ in tenants])
            keep.append((data, names))
            arr[i].kind = 1 if kind == "token" else 0
            arr[i].len = data.size
            arr[i].data = data.ctypes.data if data.size else None
            arr[i].tenants = names
            arr[i].n_tenants = len(tenants)
        applied = C.c_uint32()
        self._h.call("smgx_tree_apply_repair_page", (model_id or "").encode(), 1 if tree_kind == "token" else 0, C.cast(arr, C.c_void_p), len(entries), C.byref(applied))
        return applied.value

    def set_load_feedback(self, enabled: bool):
        """Event-driven batches as a request STREAM: every pick bumps its worker's load before the next request is decided
        (WorkerLoadGuard, routers/http/router.rs:319-321); off = one frozen fleet snapshot per batch."""
        self._h.call("smgx_set_load_feedback", 1 if enabled else 0)

    def set_tree_batch_mode(self, mode: str):
        self._h.call("smgx_set_tree_batch_mode", TREE_BATCH_MODES[mode])

    # -- tokenizer + text-in pick (the whole hot path on the device) ------------------------------------------------
    def load_tiktoken_tokenizer(self, path: str, special_tokens=None, model: str = UNKNOWN_MODEL_ID) -> TiktokenTokenizer:
        return TiktokenTokenizer(self._h, model, path, special_tokens)

    def load_hf_tokenizer(self, path: str, model: str = UNKNOWN_MODEL_ID) -> HuggingFaceTokenizer:
        return HuggingFaceTokenizer(self._h, model, path)

    def select_worker_batch_text(self, workers: Sequence[BasicWorker], texts, want_tokens: bool = True):
        """Chat-template-rendered texts → (worker_idx, info, token lists): tokenise on the GPU, then the cache-aware pick,
        without the tokens visiting the host in between (preparation.rs:134 + worker_selection.rs:157 in one call)."""
        model = self._push_fleet(workers)
        data, offsets = TiktokenTokenizer._ragged(texts)
        n = len(texts)
        out = np.full(max(n, 1), -1, dtype=np.int32)
        info = (_lib.DecisionInfo * max(n, 1))()
        cap = int(offsets[-1]) + 1
        toks, toff = (np.zeros(cap, np.uint32), np.zeros(n + 1, np.uint32)) if want_tokens else (None, None)
        self._h.call("smgx_select_batch_text", model, _p(data), _p(offsets), n, _p(out), C.cast(info, C.c_void_p),
                     _p(toks) if want_tokens else None, _p(toff) if want_tokens else None, cap)
        tokens = [toks[toff[i]:toff[i + 1]].tolist() for i in range(n)] if want_tokens else None
        return out[:n], [info[i] for i in range(n)], tokens

    def evict_cache(self, max_size: int):  # cache_aware.rs:311-352
        self._h.call("smgx_evict_cache", max_size)

    def token_tree(self, model: str = UNKNOWN_MODEL_ID) -> TokenTree:
        t = TokenTree.__new__(TokenTree)
        t.h, t.model = self._h, model.encode()
        return t

    def string_tree(self, model: str = UNKNOWN_MODEL_ID) -> "Tree":
        return Tree(self._h, model)

    def take_processed(self, model: str = UNKNOWN_MODEL_ID, n: Optional[int] = None):
        n = n if n is not None else len(self._slices.get(model, ()))
        out = np.zeros(max(n, 1), dtype=np.uint64)
        self._h.call("smgx_take_processed", model.encode(), _p(out), n)
        return out[:n]

    def kernel_launches(self) -> int:
        return _lib.load().smgx_kernel_launches(self._h.p)


# ---- adjacent policy on the same plumbing: prefix_hash ---------------------------------------------------------
PREFIX_BRANCHES = ["no_healthy_workers", "no_tokens", "ring_hit", "load_balance_walk", "fallback_least_load"]   # prefix_hash.rs:70-83


@dataclass
class PrefixHashConfig:  # policies/prefix_hash.rs:38-58
    prefix_token_count: int = 256
    load_factor: float = 1.25


class HashRing:
    """worker::HashRing (model_gateway/src/worker/hash_ring.rs): 150 virtual nodes per URL at blake3("{url}#{vnode}")[..8], hashed on
    the GPU and kept sorted inside the library.  Bound to (policy handle, model) like the registry's per-model ring."""
    _seq = 0

    def __init__(self, urls, handle: Optional[_Handle] = None, model: Optional[str] = None, device_id: int = 0):
        self._h = handle or _Handle(CacheAwareConfig(eviction_interval_secs=0), device_id)
        if model is None:
            HashRing._seq += 1
            model = f"__ring_{HashRing._seq}"
        self.model = normalize_model_key(model).encode()
        self.urls = [str(u) for u in urls]
        enc = [u.encode() for u in self.urls]
        arr = (C.c_char_p * max(len(enc), 1))(*enc)
        self._h.call("smgx_hash_ring_set", self.model, arr, len(enc))

    def entries(self):
        n = C.c_uint32(0)
        self._h.call("smgx_hash_ring_entries", self.model, None, None, 0, C.byref(n))
        pos, url = np.zeros(max(n.value, 1), np.uint64), np.zeros(max(n.value, 1), np.uint32)
        self._h.call("smgx_hash_ring_entries", self.model, _p(pos), _p(url), n.value, C.byref(n))
        return pos[:n.value], url[:n.value]

    def __len__(self):  # HashRing::len (:142-144)
        n = C.c_uint32(0)
        self._h.call("smgx_hash_ring_entries", self.model, None, None, 0, C.byref(n))
        return n.value

    def is_empty(self):
        return len(self) == 0

    def worker_count(self):  # :147-149
        return len(self) // 150

    def find_healthy_urls(self, keys, is_healthy):
        """find_healthy_url (:102-134) for a batch of keys on the GPU → list of URL or None."""
        data, offsets = TiktokenTokenizer._ragged(list(keys))
        ok = np.ascontiguousarray(np.asarray([1 if is_healthy(u) else 0 for u in self.urls] or [0], dtype=np.uint8))
        n = offsets.size - 1
        out = np.full(max(n, 1), -1, np.int32)
        self._h.call("smgx_hash_ring_find_healthy", self.model, _p(data), _p(offsets), n, _p(ok), _p(out))
        return [None if i < 0 else self.urls[i] for i in out[:n]]

    def find_healthy_url(self, key: str, is_healthy):
        return self.find_healthy_urls([key], is_healthy)[0]


class PrefixHashPolicy:
    """policies::PrefixHashPolicy (prefix_hash.rs:87-235) on the GPU: XXH3 of the first N tokens → ring lookup → bounded-load check."""

    def __init__(self, config: Optional[PrefixHashConfig] = None, device_id: int = 0, max_batch: int = 0):
        self.config = config or PrefixHashConfig()
        self._h = _Handle(CacheAwareConfig(eviction_interval_secs=0), device_id, max_batch)
        self._h.call("smgx_prefix_hash_configure", int(self.config.prefix_token_count), float(self.config.load_factor))
        self._slices = {}
        self._ring_of = {}

    @classmethod
    def with_defaults(cls, **kw):  # :100-103
        return cls(PrefixHashConfig(), **kw)

    def name(self) -> str:  # :231-233
        return "prefix_hash"

    def needs_request_text(self) -> bool:  # trait default (policies/mod.rs:77-79)
        return False

    def on_request_complete(self, worker_url: str, success: bool):  # trait default: no state
        return None

    def hash_ring(self, urls, model: str = UNKNOWN_MODEL_ID) -> HashRing:
        """The registry's per-model ring (worker/registry.rs get_hash_ring), bound to this policy's device state."""
        ring = HashRing(urls, handle=self._h, model=model)
        self._ring_of[ring.model] = ring
        return ring

    def compute_prefix_hashes(self, requests):  # compute_prefix_hash (:106-113) for a batch
        tokens, offsets = _ragged_tokens(requests)
        n = offsets.size - 1
        out = np.zeros(max(n, 1), np.uint64)
        self._h.call("smgx_prefix_hashes", _tok_ptr(tokens), _p(offsets), n, _p(out))
        return out[:n]

    def _push_fleet(self, workers, ring: Optional[HashRing]) -> bytes:
        # the ring is per model in the reference (SelectWorkerInfo.hash_ring comes from the registry keyed by model id); a policy
        # without a ring for the call uses the model's own state with the ring cleared
        model = ring.model if ring is not None else normalize_model_key(workers[0].model_id() if workers else "").encode()
        if ring is not None and ring._h is not self._h:
            raise ValueError("hash ring belongs to another policy handle: build it with PrefixHashPolicy.hash_ring()")
        if ring is None and model in self._ring_of:
            self._h.call("smgx_hash_ring_clear", model)
            del self._ring_of[model]
        urls = tuple(w.url() for w in workers)
        if self._slices.get(model) != urls:
            enc = [u.encode() for u in urls]
            arr = (C.c_char_p * max(len(enc), 1))(*enc)
            self._h.call("smgx_set_workers", model, arr, len(enc))
            self._slices[model] = urls
        loads = _u64([w.load() for w in workers])
        healthy = np.ascontiguousarray(np.asarray([1 if w.is_healthy() else 0 for w in workers] or [0], dtype=np.uint8))
        self._h.call("smgx_set_fleet_state", model, _p(loads) if len(workers) else None, _p(healthy), None, len(workers))
        return model

    def select_worker_batch(self, workers, requests=None, ring: Optional[HashRing] = None, tokens=None, offsets=None, has_tokens=None):
        """One batch against one fleet snapshot → (idx int32[n] with -1 = None, branch names)."""
        model = self._push_fleet(workers, ring)
        if requests is not None:
            has_tokens = [0 if r is None else 1 for r in requests]
            tokens, offsets = _ragged_tokens([[] if r is None else r for r in requests])
            if all(has_tokens):
                has_tokens = None
        tokens, offsets = _u32(tokens), _u32(offsets)
        n = offsets.size - 1
        out = np.full(max(n, 1), -1, np.int32)
        info = (_lib.DecisionInfo * max(n, 1))()
        flags = None if has_tokens is None else np.ascontiguousarray(np.asarray(has_tokens, np.uint8))
        self._h.call("smgx_prefix_hash_select_batch_tokens", model, _tok_ptr(tokens), _p(offsets), n, _p(flags) if flags is not None else None, _p(out),
                     C.cast(info, C.c_void_p))
        return out[:n], [PREFIX_BRANCHES[info[i].branch] for i in range(n)]

    def select_worker_impl(self, workers, info: SelectWorkerInfo):  # :203-222 → (Option<usize>, Branch)
        idx, br = self.select_worker_batch(workers, [info.tokens], ring=info.hash_ring)
        return (None if idx[0] < 0 else int(idx[0])), br[0]

    def select_worker(self, workers, info: SelectWorkerInfo) -> Optional[int]:  # LoadBalancingPolicy::select_worker (:225-229)
        return self.select_worker_impl(workers, info)[0]

    def kernel_launches(self) -> int:
        return int(self._h.L.smgx_kernel_launches(self._h.p))


def _ragged_tokens(requests):
    lens = [len(r) for r in requests]
    offsets = np.zeros(len(requests) + 1, dtype=np.uint32)
    np.cumsum(lens, out=offsets[1:])
    tokens = _u32(np.concatenate([np.asarray(r, dtype=np.uint32) for r in requests]) if sum(lens) else [])
    return tokens, offsets


def _tok_ptr(tokens):
    return _p(tokens) if tokens.size else C.cast(C.create_string_buffer(4), C.c_void_p)


class PowerOfTwoPolicy:
    """policies::PowerOfTwoPolicy (power_of_two.rs:18-135) on the GPU: two random healthy candidates per request, the less loaded one wins
    (token usage when both have a cached load response, request counts otherwise).  The reference draws from a thread-local generator;
    here every call consumes one value of a per-policy seed sequence and request i of the call uses draws 2i, 2i + 1 of that stream."""

    def __init__(self, device_id: int = 0, max_batch: int = 0, seed: int = 0x5EED):
        self._h = _Handle(CacheAwareConfig(eviction_interval_secs=0), device_id, max_batch)
        self._slices = {}
        self._seed = int(seed) & (2**64 - 1)
        self._calls = 0

    def name(self) -> str:  # :122-124
        return "power_of_two"

    def needs_request_text(self) -> bool:  # trait default (policies/mod.rs:77-79)
        return False

    def update_loads(self, loads: dict):
        """update_loads (:129-135): url → WorkerLoadResponse; a value may be the response's per-DP-rank token_usage list (→ its mean,
        effective_token_usage, protocols worker.rs:1039-1044, 0.0 when empty) or that number itself."""
        if not loads:
            return
        urls = [u.encode() for u in loads]
        vals = []
        for v in loads.values():
            if isinstance(v, (list, tuple, np.ndarray)):
                vals.append(float(np.sum(np.asarray(v, np.float64)) / len(v)) if len(v) else 0.0)
            else:
                vals.append(float(v))
        arr = (C.c_char_p * len(urls))(*urls)
        usage = np.ascontiguousarray(np.asarray(vals, np.float64))
        self._h.call("smgx_power_of_two_update_loads", arr, _p(usage), len(urls))

    def _push_fleet(self, workers) -> bytes:
        model = normalize_model_key(workers[0].model_id() if workers else "").encode()
        urls = tuple(w.url() for w in workers)
        if self._slices.get(model) != urls:
            enc = [u.encode() for u in urls]
            arr = (C.c_char_p * max(len(enc), 1))(*enc)
            self._h.call("smgx_set_workers", model, arr, len(enc))
            self._slices[model] = urls
        loads = _u64([w.load() for w in workers])
        healthy = np.ascontiguousarray(np.asarray([1 if w.is_healthy() else 0 for w in workers] or [0], dtype=np.uint8))
        circuit = np.ascontiguousarray(np.asarray([1 if w.circuit_breaker_can_execute() else 0 for w in workers] or [0], dtype=np.uint8))
        self._h.call("smgx_set_fleet_state", model, _p(loads) if len(workers) else None, _p(healthy), _p(circuit), len(workers))
        return model

    def next_seed(self) -> int:
        self._calls += 1
        return (self._seed + 0xD1B54A32D192ED03 * self._calls) & (2**64 - 1)

    def select_worker_batch(self, workers, n: int, seed: Optional[int] = None, with_details: bool = False):
        """n requests against one fleet snapshot → idx int32[n] (-1 = None) [, pairs int32[n, 2], metric uint8[n]]."""
        model = self._push_fleet(workers)
        seed = self.next_seed() if seed is None else int(seed) & (2**64 - 1)
        out = np.full(max(n, 1), -1, np.int32)
        pairs = np.full((max(n, 1), 2), -1, np.int32)
        metric = np.zeros(max(n, 1), np.uint8)
        self._h.call("smgx_power_of_two_select_batch", model, n, seed, _p(out), _p(pairs), _p(metric) if with_details else None)
        for i, pr in zip(out[:n], pairs[:n]):
            if i >= 0 and pr[0] >= 0 and hasattr(workers[int(i)], "increment_processed"):
                workers[int(i)].increment_processed()        # :110 (not on the single-healthy-worker early return, :44-46)
        return (out[:n], pairs[:n], metric[:n]) if with_details else out[:n]

    def select_worker(self, workers, info=None) -> Optional[int]:
        """LoadBalancingPolicy::select_worker (:36-120)."""
        if not workers:
            return None
        idx = self.select_worker_batch(workers, 1)
        return None if idx[0] < 0 else int(idx[0])


class PolicyFactory:
    """policies::PolicyFactory for the one policy this library replaces (factory.rs:17-93)."""

    @staticmethod
    def create_by_name(name: str, **kw):
        if name.lower() in ("cache_aware", "cacheaware"):  # factory.rs:84
            return CacheAwarePolicy(CacheAwareConfig(), **kw)
        if name.lower() in ("prefix_hash", "prefixhash"):  # factory.rs:90
            return PrefixHashPolicy.with_defaults(**kw)
        if name.lower() in ("power_of_two", "poweroftwo"):  # factory.rs:83
            return PowerOfTwoPolicy(**kw)
        return None

    @staticmethod
    def create_from_config(cache_threshold, balance_abs_threshold, balance_rel_threshold, eviction_interval_secs, max_tree_size, block_size,
                           **kw):  # factory.rs:22-39 PolicyConfig::CacheAware{..}
        return CacheAwarePolicy(CacheAwareConfig(cache_threshold, balance_abs_threshold, balance_rel_threshold, eviction_interval_secs,
                                                 max_tree_size, block_size), **kw)
