"""CPU tests of the boundary: the C ABI library loads, exports every symbol include/cssm.h
declares, and fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
from configs import c2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cssm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cssm_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(_abi.LIB_PATH)


def test_every_declared_symbol_is_exported():
    lib = _abi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cssm.h but not exported"
    assert sorted(_abi.declared_symbols()) == syms, "python binding and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _abi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (cssm_[a-z0-9_]+)", out))
    assert exported == set(syms), exported ^ set(syms)


def test_library_is_sm100a_cuda():
    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version_and_error_string():
    lib = _abi.lib()
    assert lib.cssm_version() == 100
    assert lib.cssm_last_error() is not None


def test_no_cpu_fallback_without_a_device():
    """Without a GPU every compute entry point must fail with an error, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_abi.CssmError) as e:
        cs.GpuFilterHandle(c2(), _abi.RESAMPLE_SYSTEMATIC, 128)
    assert e.value.status == -2
    with pytest.raises(_abi.CssmError):
        cs.resampling.ancestors(_abi.RESAMPLE_SYSTEMATIC, np.ones(8), np.array([0.5]))


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = _abi.lib()
    h = C.c_void_p()
    desc, keep = c2().desc()
    assert lib.cssm_filter_create(C.byref(desc), 0, 0, 0, 0, 1, 0, C.byref(h)) == -1      # no particles
    assert lib.cssm_filter_create(C.byref(desc), 10, 7, 0, 0, 1, 0, C.byref(h)) == -1     # unknown resampler
    assert lib.cssm_filter_create(None, 10, 0, 0, 0, 1, 0, C.byref(h)) == -1
    assert b"resample" in lib.cssm_last_error() or b"model" in lib.cssm_last_error()
    assert lib.cssm_filter_destroy(None) == 0
    assert lib.cssm_filter_step(None, 0.0, 1, 0.0, None, None) == -1
