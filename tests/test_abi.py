"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol
include/cueball_b200.h declares, and refuses to work without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "node-cueball_b200", "libcueball_b200.so")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not os.path.exists(SO):
        g.build()
    return SO


def declared_symbols():
    with open(os.path.join(ROOT, "include", "cueball_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_binding_covers_header(built):
    import cueball_b200 as cb
    assert sorted(cb._native.SYMBOLS) == declared_symbols()
    assert cb._native.lib().cb_abi_version() == 1


def test_struct_layouts_match_header(built):
    import cueball_b200 as cb
    assert ctypes.sizeof(cb._native.CodelState) == 48
    assert ctypes.sizeof(cb._native.BatchStats) == 10 * 8 + 48 + 16
    assert ctypes.sizeof(cb._native.Backend) == 8


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cueball_b200 as cb
    with pytest.raises(cb.CueballError) as e:
        cb.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path may not touch oracle/ (only tests, smoke() and bench baselines may)."""
    pkg = os.path.join(ROOT, "node-cueball_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.replace("no reference oracle", ""), fn


def test_every_exported_entry_point_is_placed_in_integration_md():
    """INTEGRATION.md's table says which reference interface each C-ABI entry point replaces: none may be missing."""
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        doc = f.read()
    missing = [n for n in declared_symbols() if not re.search(r"\b%s\b" % n, doc)]
    assert not missing, missing
