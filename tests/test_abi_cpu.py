"""CPU checks of the drop-in boundary: the shared library loads and exports every
symbol include/tombo_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(REPO, 'include', 'tombo_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tb2_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_declared_symbols():
    from tombo_b200 import _lib
    lib = _lib.load()
    assert lib.tb2_abi_version() == 1
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_status_messages_match_reference_strings():
    from tombo_b200 import _lib
    import oracle
    for st in list(range(0, 22)):
        assert _lib.status_message(st) == oracle.status_message(st)


def test_no_cpu_fallback_without_device():
    from tombo_b200 import _lib
    if _lib.load().tb2_device_count() > 0:
        pytest.skip('a device is present')
    with pytest.raises(_lib.TomboB200Error):
        _lib.Context(0)
