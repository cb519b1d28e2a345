"""CPU tests of the boundary: the shared library loads, exports every symbol the header declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from soundsym_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "soundsym_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "header declares %s but the library does not export it" % n
    assert sorted(_lib.SYMBOLS) == names


def test_host_only_entry_points():
    L = _lib.load()
    assert b"sm_100a" in L.ss_version()
    out = C.c_size_t()
    for n, f in ((0, 0), (1023, 0), (1024, 1), (1279, 1), (1280, 2), (5120, 17), (507150, 1978), (1163214, 4540)):
        assert L.ss_frame_count(n, C.byref(out)) == 0 and out.value == f
    assert L.ss_frame_count(10, None) == _lib.SS_ERR_INVALID


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.load()
    h = C.c_void_p()
    rc = L.ss_ctx_create(0, C.byref(h))
    assert rc == _lib.SS_ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.ss_last_error(None)
    from soundsym_b200 import api
    with pytest.raises(_lib.SoundsymError):
        api.Context(0)


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "soundsym_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("oracle/ASSUMPTIONS.h", "").replace("the oracle", "").replace("oracle's", ""), f


def test_synthetic_generators_are_seeded():
    from soundsym_b200 import synth
    a, ao = synth.segments(50, 13, seed=1)
    b, bo = synth.segments(50, 13, seed=1)
    assert np.array_equal(a, b) and np.array_equal(ao, bo) and a.shape[1] == 13
    assert ao[-1] == a.shape[0] and np.all(np.diff(ao) >= 4) and np.all(np.diff(ao) <= 32)
    s = synth.audio(0.25, seed=3)
    assert len(s) == 11025 and np.abs(s).max() <= 1.0 and np.array_equal(s, synth.audio(0.25, seed=3))
