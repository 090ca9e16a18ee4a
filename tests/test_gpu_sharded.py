"""Worker-id-sharded pick (BASELINE config 4 shape, scaled down): 2 and 4 ranks, each owning a contiguous range of the fleet,
candidates exchanged with a gloo all-gather AND through the peer-memory exchange (CUDA IPC mappings, stores + flags, no collective),
merged on every rank — picks must equal the oracle run on the WHOLE fleet.
Ranks share cuda:0 when the box has fewer GPUs than ranks (NCCL would refuse that; the exchange is plumbing)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_workers, seed, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import orc
    from smg_b200 import CacheAwareConfig, synth
    from smg_b200.sharding import ShardedEventRouter
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ndev = torch.cuda.device_count()
    bs, T, jump, B = 16, 256, 8, 300
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=bs)
    urls = synth.worker_urls(n_workers)
    router = ShardedEventRouter(urls, rank, world, CacheAwareConfig(eviction_interval_secs=0, **cfg), jump_size=jump,
                                device_id=rank % max(ndev, 1), max_tokens_per_request=T)
    rng = np.random.default_rng(seed)           # same stream on every rank
    seqs = synth.gen_sequences(150, T, seed)
    op = oix = None
    if rank == 0:
        op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
        op.set_workers(urls)
        oix = orc.PositionalIndexer(jump)
        op.attach_indexer("unknown", oix)
        op.set_kv_event_monitor(True)
        for u in urls:
            oix.intern_worker(u)
    seq_id = 1
    P = T // bs
    for s in range(len(seqs)):
        hs = orc.compute_request_content_hashes(seqs[s], bs)
        for g in rng.choice(n_workers, size=int(rng.integers(1, 6)), replace=False):
            depth = int(rng.integers(1, P + 1))
            blocks = [(seq_id + i, hs[i]) for i in range(depth)]
            seq_id += depth
            if router.owns(int(g)):
                router.indexer.apply_stored(router.local_id(int(g)), blocks)
            if rank == 0:
                oix.apply_stored(int(g), blocks)
    def all_gather(arr):
        t = torch.from_numpy(arr.copy())
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return np.concatenate([o.numpy() for o in outs])
    router.connect_peers(all_gather)          # CUDA IPC handles travel over the same gloo channel, once
    ok = True
    for round_no in range(3):
        loads = synth.poisson_loads(n_workers, 8, seed + round_no)
        healthy = (rng.random(n_workers) > 0.1).astype(np.uint8)
        if round_no == 2:
            loads[n_workers // 2] += 500      # imbalanced fleet → every pick is the global first-min-load worker
        router.set_fleet_state(loads, healthy)
        q = synth.gen_queries(seqs, B, seed + round_no, block=bs)
        lens = rng.integers(0, T + 1, size=B); lens[: B // 3] = T
        reqs = [q[i, : lens[i]] for i in range(B)]
        flat = np.concatenate(reqs).astype(np.uint32)
        offs = np.zeros(B + 1, np.uint32); np.cumsum(lens, out=offs[1:])
        got = router.select(flat, offs, T, all_gather)
        fused = router.select_fused(flat, offs, T)           # peer-memory exchange: same picks on every rank
        ok = ok and np.array_equal(got, fused)
        if rank == 0:
            op.set_state(loads, healthy, [1] * n_workers)
            want, br, _, _ = op.select_batch_tokens(flat, offs.astype(np.uint64))
            ok = ok and np.array_equal(got, want)
            if round_no < 2:
                ok = ok and (np.asarray(br) == 2).sum() > B // 10
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_workers,seed", [(2, 128, 1), (2, 200, 2), (4, 300, 3),
                                                   (8, 4096, 4)])   # BASELINE config 4 shape: 8 ranks × 512 workers (ranks share the device)
def test_sharded_pick_matches_oracle_on_whole_fleet(world, n_workers, seed):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + seed
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_workers, seed, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs)
    assert all(ret.get(r) is True for r in range(world))
