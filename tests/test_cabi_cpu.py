"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from gparml_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "gparml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gparml_[a-z_0-9A-Z]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m gparml_b200.build` (or __graft_entry__.build())"
    lib = _lib.load()
    assert lib.gparml_abi_version() == 1


def test_every_header_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    # and the ctypes prototype table covers the same set
    assert sorted(_lib.PROTOTYPES) == names


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the host."""
    lib = _lib.load()
    if lib.gparml_device_count() > 0:
        pytest.skip("a GPU is visible here")
    from gparml_b200.engine import ShardContext
    with pytest.raises(_lib.GparmlError) as e:
        ShardContext(4, 2, 3, 10)
    assert "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gparml_b200/ may import it."""
    pkg = os.path.join(ROOT, "gparml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "gparml_oracle" not in txt, f
