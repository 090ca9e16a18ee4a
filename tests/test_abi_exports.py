"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/smgx.h declares,
honours the error conventions, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os

import pytest

from smg_b200 import _lib


def _no_gpu():
    return not (os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert missing == []


def test_name_and_abi_version():
    L = _lib.load()
    assert L.smgx_policy_name() == b"cache_aware"  # cache_aware.rs:704-706
    assert L.smgx_abi_version() == 1


def test_default_config_matches_reference_defaults():
    c = _lib.Config()
    _lib.load().smgx_default_config(C.byref(c))
    # CacheAwareConfig::default() (policies/mod.rs:106-117)
    assert abs(c.cache_threshold - 0.5) < 1e-7 and c.balance_abs_threshold == 32 and abs(c.balance_rel_threshold - 1.1) < 1e-6
    assert c.eviction_interval_secs == 30 and c.max_tree_size == 10000 and c.block_size == 16


def test_null_arguments_are_invalid_argument_with_message():
    L = _lib.load()
    err = C.c_char_p()
    code = L.smgx_set_workers(None, b"m", None, 0, C.byref(err))
    assert code == _lib.INVALID_ARGUMENT and b"null pointer" in err.value
    L.smgx_free_string(C.cast(err, C.c_void_p))
    assert L.smgx_policy_create(None, None) is None


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    L = _lib.load()
    c = _lib.Config()
    L.smgx_default_config(C.byref(c))
    err = C.c_char_p()
    p = L.smgx_policy_create(C.byref(c), C.byref(err))
    assert not p and b"no CPU fallback" in err.value
    L.smgx_free_string(C.cast(err, C.c_void_p))


def test_host_mirror_mode_writes_but_never_selects():
    """device_id = -1: the index writers run (host mirror) so their logic is testable without a GPU; every query fails."""
    from smg_b200.policy import CacheAwareConfig, PositionalIndexer, _Handle, ApplyError
    h = _Handle(CacheAwareConfig(eviction_interval_secs=0), device_id=-1)
    ix = PositionalIndexer(h, "unknown", 64)
    w = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w, [(1, 10), (2, 20), (3, 30)])
    assert ix.current_size() == 3 and ix.entry_count() == 3
    with pytest.raises(ApplyError, match="ParentBlockNotFound"):
        ix.apply_stored(w, [(9, 90)], parent=12345)
    with pytest.raises(_lib.SmgxError) as e:
        ix.find_matches([10, 20, 30])
    assert e.value.code == _lib.DEVICE_ERROR and "no CPU fallback" in e.value.msg
