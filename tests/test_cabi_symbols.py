"""The C-ABI library loads on a CPU-only host and exports every symbol the header declares."""
import ctypes as C

from synergynet_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/synergy_b200.h but not exported'
    assert set(declared) == set(_lib.SIGNATURES), 'ctypes table and header out of sync'
    assert lib.syn_abi_version() == 1


def test_null_handle_calls_fail_cleanly():
    lib = _lib.load()
    assert lib.syn_commit(None) == 1
    assert lib.syn_forward(None, None, 1, None, None, None) == 1
    assert lib.syn_launch_count(None) == -1
    assert b'null handle' in lib.syn_last_error()
    lib.syn_destroy(None)
    h = C.c_void_p()
    rc = lib.syn_create(0, C.byref(h))      # no GPU here -> CUDA error, not a crash
    if rc != 0:
        assert rc in (1, 2, 6) and h.value is None
    else:
        lib.syn_destroy(h)
