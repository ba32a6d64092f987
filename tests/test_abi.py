"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/tqsim.h declares;
without a GPU it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tensorrl_qas_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tqsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/tqsim.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert built_lib.tq_version() == 100


def test_library_is_sm100a_only(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tensorrl_qas_b200 import Simulator, TqError
    with pytest.raises(TqError) as e:
        Simulator(4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensorrl_qas_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_exception_barrier_on_cpu(built_lib):
    """include/tqsim.h: nothing throws or aborts across the ABI.  Entry points that need no GPU are driven into C++
    exceptions (std::length_error / std::bad_alloc from absurd sizes) and must come back with an error, not terminate."""
    import ctypes
    from tensorrl_qas_b200 import _lib
    L = _lib.lib()
    ip, dp = _lib.c_int_p, _lib.c_dbl_p
    z = (ctypes.c_int32 * 1)(0)
    f = (ctypes.c_double * 1)(0.0)
    ptr = L.tq_plan_dump(4, -1, ctypes.cast(z, ip), ctypes.cast(z, ip), ctypes.cast(z, ip), ctypes.cast(z, ip),
                         ctypes.cast(f, dp), 0, 12, 3, 0, None)
    assert ptr
    text = ctypes.string_at(ptr).decode()
    L.tq_free(ptr)
    assert text.startswith("ERROR")
    h = ctypes.c_void_p()
    x0 = (ctypes.c_double * 1)(0.0)
    rc = L.tq_cobyla_create(2 ** 31 - 1, ctypes.cast(x0, dp), 1.0, 1e-4, 10, ctypes.byref(h))
    assert rc in (-4, -1) and not h.value      # TQ_ENOMEM (or TQ_EINVAL): reported, not thrown
