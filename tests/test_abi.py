"""The C-ABI library loads and exports every symbol include/scb200.h declares;
the product path has no CPU fallback and does not import the oracle."""
import ast
import ctypes
import glob
import os
import re
from os.path import join

import numpy as np
import pytest

from .conftest import ROOT, has_cuda


def declared_symbols():
    text = open(join(ROOT, "include", "scb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from springcraft_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (make / __graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(handle, name), f"{name} is declared in include/scb200.h but not exported"
    # the ctypes binding covers exactly the declared surface
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_status_strings():
    from springcraft_b200 import _lib
    h = _lib.lib()
    assert h.scb_version() == 100
    assert h.scb_status_string(0) == b"ok"
    assert b"TabulatedForceField" in h.scb_status_string(-3)
    assert h.scb_launch_count() >= 0
    # size helpers are pure host functions (no GPU needed)
    assert h.scb_scan_scratch_bytes(10_000) > 0
    assert h.scb_eig_lowest_workspace_bytes(3, 2, 300, 32, 6, 24000) > 2 * 3 * 900 * 32 * 8
    assert h.scb_dcc_workspace_bytes(3, 1000, 50) >= 2 * 1000 * 150 * 8
    assert h.scb_eig_full_workspace_bytes(1, 60) > 0 and h.scb_eig_full_workspace_bytes(1, 3000) > 2 * 3008 ** 2 * 8


def test_status_to_exception_mapping():
    from springcraft_b200 import _lib
    for status, exc in ((-1, ValueError), (-3, ValueError), (-6, NotImplementedError), (-5, RuntimeError),
                        (-4, RuntimeError)):
        with pytest.raises(exc):
            _lib.check(status)
    assert _lib.check(0) == 0
    assert _lib.check(-5, allow=(-5,)) == -5


def test_struct_layouts_match_header():
    from springcraft_b200 import _lib
    # 4 x int32 + double + 11 pointers
    assert ctypes.sizeof(_lib.FFDesc) == 16 + 8 + 11 * 8
    assert _lib.FFDesc.cutoff_sq.offset == 16 and _lib.FFDesc.bonded.offset == 24
    assert ctypes.sizeof(_lib.Patch) == 8 + 3 * 8


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_gpu():
    import springcraft_b200 as sc
    coord = np.random.default_rng(0).random((10, 3)) * 10
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sc.compute_kirchhoff(coord, sc.InvariantForceField(7.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sc.ANM(coord, sc.HinsenForceField()).eigen()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sc.enm_ensemble(coord[None], sc.InvariantForceField(7.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sc.HinsenForceField().force_constant(np.array([0]), np.array([1]), np.array([9.0]))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under springcraft_b200/ may import it."""
    for path in glob.glob(join(ROOT, "springcraft_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), path
    for path in glob.glob(join(ROOT, "springcraft_b200", "csrc", "*")):
        assert "oracle" not in open(path).read(), path
