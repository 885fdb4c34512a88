"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the
public header declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C

import pytest

from conftest import have_gpu


def test_library_exports_every_declared_symbol():
    from incompact3d_b200 import _lib
    L = _lib.load()
    names = _lib.symbols_declared_in_header()
    assert len(names) >= 85
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from incompact3d_b200 import X3D, X3DError
    with pytest.raises(X3DError, match="no CUDA device"):
        X3D(0)


def test_product_does_not_reference_oracle():
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "incompact3d_b200", "**", "*"), recursive=True):
        if os.path.isfile(f) and f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")):
            txt = open(f, errors="ignore").read()
            assert "x3d_oracle" not in txt and "oracle_lib" not in txt and "libx3d_oracle" not in txt, f
