"""Stress the event-driven pick at config-2 scale: many launches of different shapes, every pick compared with the oracle's, mismatches
printed with their position and kind.  python tools/stress_stream.py [reps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from smg_b200 import _lib  # noqa: E402
from tests.test_gpu_scale import _config2  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_seq, W, T, bs, B, NB = 31250, 64, 512, 16, 4096, 37
    pol, ws, ix, op, seqs = _config2(n_seq, W, T, bs, B)
    h, L = pol._h, _lib.load()
    model = pol._push_fleet(ws)
    err = _lib.new_err()
    offsets = (np.arange(B + 1, dtype=np.uint64) * T).astype(np.uint32)
    d_off = L.smgx_device_alloc(h.p, offsets.nbytes, C.byref(err))
    h.call("smgx_memcpy_h2d", d_off, offsets.ctypes.data_as(C.c_void_p), offsets.nbytes)
    host, d_tok, d_out, want = [], [], [], []
    off64 = offsets.astype(np.uint64)
    for r in range(NB):
        flat = np.ascontiguousarray(bench.gen_batch(seqs, B, 900 + r, bs)[0].reshape(-1))
        host.append(flat)
        dt = L.smgx_device_alloc(h.p, flat.nbytes, C.byref(err))
        h.call("smgx_memcpy_h2d", dt, flat.ctypes.data_as(C.c_void_p), flat.nbytes)
        d_tok.append(dt)
        d_out.append(L.smgx_device_alloc(h.p, B * 4, C.byref(err)))
        want.append(np.asarray(op.select_batch_tokens(flat, off64)[0]))
    got = np.zeros(B, np.int32)
    poison = np.full(B, -7, np.int32)
    total_bad = 0
    mode = sys.argv[2] if len(sys.argv) > 2 else "mix"
    if mode == "k32":
        shapes = [list(range(32)), list(range(5, 37))]
    elif mode == "singles":
        shapes = [[i] for i in range(NB)]
    else:
        shapes = [[i] for i in range(0, NB, 5)] + [list(range(20)), list(range(5, 37)), list(range(37)), list(range(32))]
    for rep in range(reps):
        for shape in shapes:
            for j in shape:
                h.call("smgx_memcpy_h2d", d_out[j], poison.ctypes.data_as(C.c_void_p), B * 4)
            n = len(shape)
            h.call("smgx_select_many_tokens_device", model, n, (C.c_void_p * n)(*[d_tok[j] for j in shape]), (C.c_void_p * n)(*[d_off] * n),
                   (C.c_uint32 * n)(*[B] * n), T, (C.c_void_p * n)(*[d_out[j] for j in shape]))
            h.call("smgx_synchronize")
            for pos, j in enumerate(shape):
                h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[j], B * 4)
                bad = np.nonzero(got != want[j])[0]
                for r in bad[:8]:
                    print(f"rep {rep} shape n={n} batch#{pos} (ring {j}) request {r}: got {got[r]} want {want[j][r]}", flush=True)
                total_bad += len(bad)
    print(f"stress: {reps} reps x {len(shapes)} shapes, {total_bad} mismatches", flush=True)
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
