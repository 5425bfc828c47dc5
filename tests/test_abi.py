"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol the header declares.
No compute call is made here (no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rsoccer_b200 import _lib
    _lib.build()
    return _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "rsoccer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "librsoccer_b200.so does not export " + n
    assert sorted(lib.SYMBOLS) == names, "rsoccer_b200/_lib.py SYMBOLS out of sync with the header"


def test_version_and_error_string(lib):
    L = lib.lib()
    assert L.rs_version() >= 100
    assert isinstance(L.rs_last_error(), (bytes, type(None)))


def test_invalid_arguments_are_rejected_before_touching_the_gpu(lib):
    L = lib.lib()
    h = ctypes.c_void_p()
    assert L.rs_create(0, 0, 3, 3, 25, 0, -1, 0, 0, ctypes.byref(h)) == -1          # n_envs < 1
    assert b"n_envs" in L.rs_last_error()
    assert L.rs_create(7, 0, 3, 3, 25, 4, -1, 0, 0, ctypes.byref(h)) == -1          # unknown kind
    assert L.rs_create(0, 9, 3, 3, 25, 4, -1, 0, 0, ctypes.byref(h)) == -1          # unknown field
    assert L.rs_create(0, 0, 30, 3, 25, 4, -1, 0, 0, ctypes.byref(h)) == -1         # too many robots
    assert L.rs_create(0, 0, 3, 3, 25, 4, -1, 0, 2 ** 32, ctypes.byref(h)) == -1    # env ids must fit 32 bits
    assert not h.value


def test_no_cpu_fallback(lib):
    """without a CUDA device the product fails loudly (and never routes through oracle/)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib.lib()
    h = ctypes.c_void_p()
    assert L.rs_create(0, 0, 3, 3, 25, 4, -1, 0, 0, ctypes.byref(h)) == -2           # RS_E_CUDA
    assert b"no CUDA device" in L.rs_last_error()
    from rsoccer_b200 import engine
    with pytest.raises(lib.RsError):
        engine.BatchedWorld(0, 0, 3, 3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rsoccer_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.lower().replace("the oracle", "").replace("oracle/", "").replace(
                    "oracle's", "").replace("oracle of", "") or "import" not in txt or all(
                    "oracle" not in ln for ln in txt.splitlines() if ln.strip().startswith(("import", "from", "#include"))), f


def test_argument_validation_needs_no_device(lib):
    """bad sizes / unknown worlds are rejected before any CUDA call; rs_task_act_dim is a pure
    table of the reference action spaces (vss_gym.py:64, static_defenders.py:54,
    dribbling.py:50-51, pass_endurance.py:53)."""
    L = lib.lib()
    h = ctypes.c_void_p()
    for args in ((0, 0, 3, 3, 25, 0), (0, 0, 3, 3, 0, 4), (0, 7, 3, 3, 25, 4), (1, 2, 12, 11, 25, 4), (2, 0, 1, 1, 25, 4)):
        assert L.rs_create(*args, -1, 0, 0, ctypes.byref(h)) == -1 and not h.value       # RS_E_INVALID
    assert L.rs_create(0, 0, 3, 3, 25, 4, -1, 0, 2 ** 32, ctypes.byref(h)) == -1          # env ids must fit 32 bits
    assert [L.rs_task_act_dim(t) for t in range(5)] == [2, 5, 5, 4, 3]
    assert L.rs_task_act_dim(9) < 0 and b"unknown task" in L.rs_last_error()
