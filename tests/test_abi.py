"""CPU: the C-ABI library loads and exports every symbol include/xb200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "xb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from xuance_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xb200.h but not exported by libxb200.so"
    assert sorted(_lib.exported_symbols()) == names, "python binding table and header disagree"
    assert lib.xb_version() >= 100


def test_error_strings_and_argument_checks_without_gpu():
    from xuance_b200 import _lib
    lib = _lib.load()
    assert b"aligned" in lib.xb_error_string(-2)
    # argument validation happens before any CUDA call: NULL pointers are rejected with XB_EINVAL
    assert lib.xb_gae_scan(None, None, None, None, None, None, None, None, 4, 4, 0.99, 0.95, 1, None) == -1
    assert lib.xb_rollout_store(None, None, 16, None, None, 0, 0, 4, 0, None) == -1
    assert lib.xb_per_insert(None, None, None, 1, 3, 0, 0.5, None) == -1


def test_product_refuses_cpu():
    """No CPU fallback: buffers / learners raise on a non-CUDA device instead of silently running elsewhere."""
    import numpy as np
    import pytest
    from xuance_b200.common import DummyOnPolicyBuffer, Box, Discrete
    with pytest.raises(RuntimeError, match="CUDA"):
        DummyOnPolicyBuffer(Box(-1, 1, (3,), np.float32), Discrete(2), None, 2, 4, device="cpu")


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under xuance_b200/ may import it."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "xuance_b200")):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
