"""CPU: the C-ABI library builds, loads and exports every symbol include/cobel_b200.h declares
(no compute call is made -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    return ctypes.CDLL(g.LIB)


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'cobel_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cobel_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built):
    names = declared_functions()
    assert 'cobel_dynaq_run' in names
    for n in names:
        assert hasattr(built, n), 'missing export %s' % n


def test_binding_covers_header(built):
    from cobel_rl_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared_functions()
    assert built.cobel_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match(built):
    """ctypes mirrors must have the C layout: compare against sizes the library reports."""
    from cobel_rl_b200 import _lib
    built.cobel_sizeof.restype = ctypes.c_size_t
    built.cobel_sizeof.argtypes = [ctypes.c_char_p]
    for name, cls in _lib.STRUCTS.items():
        assert built.cobel_sizeof(name.encode()) == ctypes.sizeof(cls), name


def test_no_fallback_without_library(monkeypatch):
    from cobel_rl_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libcobel_b200.so')
    with pytest.raises(_lib.CobelError):
        _lib.lib()


def test_integration_md_stub_matches_library(built, monkeypatch):
    """The ctypes stub shown in INTEGRATION.md must mirror the header: execute its struct definitions
    against the built library (its own cobel_sizeof assertion runs)."""
    import __graft_entry__ as g
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text.split('```python')[2].split('```')[0]
    defs = block.split('def dynaq_train_b200')[0]
    monkeypatch.setenv('COBEL_B200_LIB', g.LIB)
    ns = {}
    exec(compile(defs, 'INTEGRATION.md', 'exec'), ns)
    assert ctypes.sizeof(ns['DynaQParams']) == built.cobel_sizeof(b'CobelDynaQParams')
