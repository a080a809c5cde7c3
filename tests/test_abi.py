"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/next3d_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'next3d_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(n3d_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, 'next3d_b200', 'libnext3d_b200.so')
    assert os.path.exists(so), 'build the library first: make -C next3d_b200/csrc'
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    lib.n3d_version.restype = ctypes.c_int
    assert lib.n3d_version() >= 100


def test_python_binding_covers_header():
    from next3d_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_invalid_arguments_are_reported_not_crashing():
    from next3d_b200 import _lib
    rc = _lib.lib.n3d_bias_act(None, None, None, 0, 16, 0, 0, 3, 0.2, 1.0, -1.0, None)
    assert rc == -1 and b'null' in _lib.lib.n3d_last_error()
    rc = _lib.lib.n3d_fill_mouth(None, 0, 0, 0, None)
    assert rc == -1
