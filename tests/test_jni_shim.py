"""The JVM side of the boundary without a JVM: jvm/cssm_jni.c is compiled with -Wall -Werror
against a stand-in jni.h (same names and signatures as the JDK's, jvm/jni_stub/jni.h), linked against libcssm_gpu.so,
and its Java_... entry points are driven through a fake JNIEnv by tests/jni_harness.c.  The Scala sources
(jvm/*.scala) are checked for the natives they declare: every @native method has a Java_ entry point and vice versa."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi, Model, Sde, SdeParameter, Parameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "composablestatespacemodels_b200", "csrc")
BUILD = os.path.join(ROOT, "oracle", "_build")


def build_harness():
    import __graft_entry__ as g
    g.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "jni_harness")
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "jvm", "jni_stub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "jvm", "cssm_jni.c"), os.path.join(ROOT, "tests", "jni_harness.c"), "-L" + CSRC, "-lcssm_gpu",
           "-Wl,-rpath," + CSRC, "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return exe


def test_shim_compiles_warning_free_and_rethrows_library_errors():
    exe = build_harness()
    out = subprocess.run([exe, "errors"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert 'throws 1 msg "total particle count must be in [1, 2^31-1]"' in lines[0] and "handle 0" in lines[0]
    assert 'throws 2 msg "null filter handle"' in lines[1]
    assert 'throws 3 msg "null filter handle"' in lines[2]
    assert lines[3].endswith("throws 3")                       # destroy(0) is a no-op, as AutoCloseable.close() must be


def test_every_native_has_an_entry_point_and_the_adapter_uses_the_resident_series():
    scala = open(os.path.join(ROOT, "jvm", "CssmNative.scala")).read()
    natives = set(re.findall(r"@native def (\w+)\(", scala))
    shim = open(os.path.join(ROOT, "jvm", "cssm_jni.c")).read()
    entries = set(re.findall(r"Java_com_github_jonnylaw_gpu_CssmNative_00024_(\w+)\(", shim))
    assert natives == entries, natives ^ entries
    for n in ("filterLoadSeries", "filterLlResident", "filterSetParams", "filterSetTieRule", "filterReseed", "filterSeriesLen"):
        assert n in natives, n
    # every C ABI function the shim calls exists in the header
    header = open(os.path.join(ROOT, "include", "cssm.h")).read()
    for fn in set(re.findall(r"\b(cssm_[a-z0-9_]+)\(", shim)):
        assert re.search(r"\b%s\(" % fn, header), fn
    # INTEGRATION.md promises ONE handle per BootstrapFilter, re-parameterised per proposal
    adapter = open(os.path.join(ROOT, "jvm", "FilterGpu.scala")).read()
    boot = adapter[adapter.index("object FilterGpu {"):]
    assert "filterSetParams" in boot and "filterLlResident" in boot and "filterLoadSeries" in boot
    assert "FilterGpu(ms, m, resampleKind)" not in boot        # no handle per proposal


@pytest.mark.gpu
def test_shim_drives_a_filter_and_agrees_with_the_ctypes_path():
    exe = build_harness()
    out = subprocess.run([exe, "filter"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j["throws"] == 0 and j["series_len"] == 40 and j["sample_is_member"] == 1
    assert j["ll1"] == j["ll1b"]                                # reseeded alike: same bits
    assert j["ll2"] != j["ll1"] and j["ll3"] == j["ll2"] and j["ll4"] == j["ll2"]   # resident, host-buffer and stepping forms agree
    assert j["anc"] == [1, 2, 3, 3]
    # the same filter through the Python mirror
    def mod(sigma):
        return Model.poisson(Sde.ouProcess(1))(Parameters(None, SdeParameter.ouParameter([1.0], [0.5], [0.2], [1.5], [sigma])))
    t = 0.1 * np.arange(40)
    y = np.array([(s * 7 + 3) % 6 for s in range(40)], dtype=np.float64)
    ho = np.ones(40, dtype=np.uint8)
    ho[5] = 0
    h = cs.GpuFilterHandle(mod(0.05), _abi.RESAMPLE_SYSTEMATIC, 4096, dtype=_abi.F64, seed=11)
    h.load_series(t, y, ho)
    assert h.ll_resident() == j["ll1"]
    h.set_params(mod(0.2))
    h.reseed(11, 0)
    assert h.ll_resident() == j["ll2"]
    h.close()
