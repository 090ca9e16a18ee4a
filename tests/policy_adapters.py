"""Adapter giving the CPU oracle the same surface as smg_b200.CacheAwarePolicy, so policy-level scenarios
(tests/scenarios_cache_aware.py) run unchanged against the oracle (CPU) and the CUDA path (GPU)."""
from oracle import orc
from smg_b200.policy import BasicWorker, CacheAwareConfig, SelectWorkerInfo, normalize_model_key  # noqa: F401


class _OracleMonitor:
    def __init__(self, pol, default_block_size=None):
        self.pol, self.indexers, self.default_block_size = pol, {}, default_block_size

    def create_indexer(self, model, jump_size=64):
        ix = orc.PositionalIndexer(jump_size)
        self.indexers[model] = ix
        self.pol._o.attach_indexer(model, ix)
        return ix

    def get_indexer(self, model):
        return self.indexers.get(model)

    def set_block_size(self, model, bs):
        self.pol._o.set_block_size(model, bs)


class OraclePolicy:
    def __init__(self, config=None):
        c = config or CacheAwareConfig()
        self.config = c
        self._o = orc.CacheAwarePolicy(c.cache_threshold, c.balance_abs_threshold, c.balance_rel_threshold, c.eviction_interval_secs,
                                       c.max_tree_size, c.block_size)
        self._urls = None
        self.last = None

    def name(self):
        return "cache_aware"

    def needs_request_text(self):
        return True

    def _set(self, workers, init):
        urls = [w.url() for w in workers]
        models = [w.model_id() for w in workers]
        if init or self._urls != (urls, models):
            self._o.set_workers(urls, models, init=init)
            self._urls = (urls, models)
        self._o.set_state([w.load() for w in workers], [w.is_healthy() for w in workers], [w.circuit_breaker_can_execute() for w in workers])

    def init_workers(self, workers):
        self._set(workers, True)

    def remove_worker_by_url(self, url):
        pass

    def kv_event_monitor(self, default_block_size=None):
        return _OracleMonitor(self, default_block_size)

    def set_kv_event_monitor(self, monitor):
        self._o.set_kv_event_monitor(monitor is not None)

    def has_event_indexer(self, model="unknown"):
        return self._o.has_event_indexer(model)

    def select_worker(self, workers, info):
        self._set(workers, False)
        d = self._o.select_worker(request_text=info.request_text, tokens=list(info.tokens) if info.tokens is not None else None)
        self.last = d
        return d.idx

    def evict_cache(self, max_size):
        self._o.evict_cache(max_size)
