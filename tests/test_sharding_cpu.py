"""CPU (gloo, world_size 2) coverage of the N>1 host logic: shard ranges, the candidate wire format and the exchange
plumbing used by smg_b200.sharding — no kernels run here."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_are_contiguous_and_balanced():
    from smg_b200.sharding import shard_range
    for n, w in ((4096, 8), (300, 4), (7, 3), (5, 8)):
        rs = [shard_range(n, g, w) for g in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1


def test_wire_structs_match_header():
    import ctypes as C

    class Cand(C.Structure):
        _fields_ = [("score", C.c_uint32), ("local_idx", C.c_uint32), ("load", C.c_uint64), ("tree_size", C.c_uint64)]

    class Fleet(C.Structure):
        _fields_ = [("min_load_idx", C.c_int32), ("first_healthy", C.c_int32), ("n_healthy", C.c_uint32), ("imbalanced", C.c_uint32),
                    ("min_load", C.c_uint64), ("max_load", C.c_uint64), ("min_healthy_load", C.c_uint64)]
    from smg_b200.sharding import CAND_BYTES, FLEET_BYTES
    assert C.sizeof(Cand) == CAND_BYTES and C.sizeof(Fleet) == FLEET_BYTES


def _gloo_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arr = np.full(48, rank + 1, np.uint8)
    t = torch.from_numpy(arr.copy())
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    got = np.concatenate([o.numpy() for o in outs])
    ret[rank] = bool((got[:48] == 1).all() and (got[48:] == 2).all())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_all_gather_rank_order_world2():
    """The merge relies on shard g's candidates landing at [g][...] — rank order of the all-gather."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29655, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs) and ret.get(0) and ret.get(1)
