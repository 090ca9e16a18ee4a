"""Launch shapes of the event-driven pick at config-2 scale with POISONED outputs: single batches, 20, 32 and 37 batches per call (one launch,
and a 32 + 5 split over two lanes), for the default pair and for the persistent streaming kernel.  Every output word must have been written
and must equal the oracle's pick — a request nobody decided shows up as the poison value, which a comparison against a previous run's
identical picks (as in the bench) cannot see.  (tools/stress_stream.py is the long-running form of the same loop.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", [0, 3])
def test_every_pick_written_and_equal_to_the_oracle(path):
    import bench
    from smg_b200 import _lib
    from tests.test_gpu_scale import _config2
    n_seq, W, T, bs, B, NB = 31250, 64, 512, 16, 4096, 37
    L = _lib.load()
    L.smgx_set_event_path(path, 0)
    try:
        pol, ws, ix, op, seqs = _config2(n_seq, W, T, bs, B)
        h = pol._h
        model = pol._push_fleet(ws)
        err = _lib.new_err()
        offsets = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
        off64 = offsets.astype(np.uint64)
        d_off = L.smgx_device_alloc(h.p, offsets.nbytes, C.byref(err))
        h.call("smgx_memcpy_h2d", d_off, offsets.ctypes.data_as(C.c_void_p), offsets.nbytes)
        d_tok, d_out, want = [], [], []
        for r in range(NB):
            flat = np.ascontiguousarray(bench.gen_batch(seqs, B, 900 + r, bs)[0].reshape(-1))
            dt = L.smgx_device_alloc(h.p, flat.nbytes, C.byref(err))
            h.call("smgx_memcpy_h2d", dt, flat.ctypes.data_as(C.c_void_p), flat.nbytes)
            d_tok.append(dt)
            d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
            want.append(np.asarray(op.select_batch_tokens(flat, off64)[0]))
        got = np.zeros(B, np.int32)
        poison = np.full(B, -7, np.int32)
        shapes = [[0], [11], [36], list(range(20)), list(range(5, 37)), list(range(37))]
        for rep in range(2):
            for shape in shapes:
                for j in shape:
                    h.call("smgx_memcpy_h2d", d_out[j], poison.ctypes.data_as(C.c_void_p), B * 4)
                n = len(shape)
                h.call("smgx_select_many_tokens_device", model, n, (C.c_void_p * n)(*[d_tok[j] for j in shape]), (C.c_void_p * n)(*[d_off] * n),
                       (C.c_uint32 * n)(*[B] * n), T, (C.c_void_p * n)(*[d_out[j] for j in shape]))
                h.call("smgx_synchronize")
                for j in shape:
                    h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[j], B * 4)
                    assert not (got == -7).any(), f"path {path}, {n} batches per call: {(got == -7).sum()} picks of ring batch {j} were never written"
                    assert np.array_equal(got, want[j]), f"path {path}, {n} batches per call, ring batch {j}: {(got != want[j]).sum()} picks differ from the oracle"
        for d in d_tok + d_out + [d_off]:
            L.smgx_device_free(h.p, d)
    finally:
        L.smgx_set_event_path(0, 0)
